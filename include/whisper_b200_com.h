// whisper_b200 — COM-style API, binary compatible with the reference's public interfaces.
//
// The reference's clients (Examples/main, WhisperNet, WhisperPS, WhisperDesktop) talk to Whisper.dll through IUnknown-compatible
// vtables and a handful of exported C functions.  This header re-declares those contracts (same GUIDs, same vtable slot order,
// same POD layouts) so that a client compiled against the reference headers can be pointed at this library:
//
//   IUnknown                 ComLightLib/unknwn.h:26-35            slots 0-2: QueryInterface, AddRef, Release
//   iContext                 Whisper/API/iContext.cl.h:23-44       {b9956374-3b18-4943-90f2-2ab18a404537}
//   iModel                   Whisper/API/iContext.cl.h:46-60       {abefb4c9-e8d8-46a3-8747-5afbadef1adb}
//   iTranscribeResult        Whisper/API/iTranscribeResult.cl.h:7-14  {2871a73f-5ce3-48f8-8779-6582ee11935e}
//   iAudioBuffer             Whisper/API/iMediaFoundation.cl.h:9-17   {013583aa-c9eb-42bc-83db-633c2c317051}
//   iAudioReader             Whisper/API/iMediaFoundation.cl.h:19-26  {35b988da-04a6-476a-a193-d8891d5dc390}  (input of iContext::runStreamed)
//   iAudioCapture            Whisper/API/iMediaFoundation.cl.h:28-34  {747752c2-d9fd-40df-8847-583c781bf013}  (input of iContext::runCapture)
//   iMediaFoundation         Whisper/API/iMediaFoundation.cl.h:36-46  {fb9763a5-d77d-4b6e-aff8-f494813cebd8}  (RIFF/WAVE media layer on Linux)
//   sCaptureParams, eCaptureStatus, sCaptureCallbacks, sCaptureDevice   Whisper/API/MfStructs.h:5-52
//   sFullParams, flags       Whisper/API/sFullParams.h:5-130
//   sSegment, sToken, ...    Whisper/API/TranscribeStructs.h:8-137
//   sModelSetup, callbacks   Whisper/API/sModelSetup.h:6-41, sLoadModelCallbacks.h:5-14, loggerApi.h:7-34, SpecialTokens.h, sLanguageList.h
//   exports                  Whisper/whisper.def:1-8: setupLogger, loadModel, findLanguageKeyW/A, getSupportedLanguages, listGPUs, initMediaFoundation
//
// Implementation: whisper_b200/csrc/com_shell.cpp — thin shells over the C ABI (include/whisper_b200.h).
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#if defined( _MSC_VER )
#define WSPCALL __stdcall
#elif defined( __i386__ )
#define WSPCALL __attribute__( ( stdcall ) )
#else
#define WSPCALL
#endif

#ifndef WSP_COM_NO_HRESULT
typedef int32_t HRESULT;
#define WSP_HR( x ) ( (HRESULT)( x ) )
constexpr HRESULT S_OK = 0, S_FALSE = 1;
constexpr HRESULT E_NOTIMPL = WSP_HR( 0x80004001 ), E_NOINTERFACE = WSP_HR( 0x80004002 ), E_POINTER = WSP_HR( 0x80004003 ), E_ABORT = WSP_HR( 0x80004004 ),
	E_FAIL = WSP_HR( 0x80004005 ), E_UNEXPECTED = WSP_HR( 0x8000FFFF ), E_OUTOFMEMORY = WSP_HR( 0x8007000E ), E_INVALIDARG = WSP_HR( 0x80070057 ),
	E_BOUNDS = WSP_HR( 0x8000000B );
inline bool SUCCEEDED( HRESULT hr ) { return hr >= 0; }
inline bool FAILED( HRESULT hr ) { return hr < 0; }

struct GUID
{
	uint32_t Data1;
	uint16_t Data2, Data3;
	uint8_t Data4[ 8 ];
	bool operator==( const GUID& o ) const { return 0 == memcmp( this, &o, sizeof( GUID ) ); }
};
typedef const GUID& REFIID;
#endif

namespace ComLight
{
	struct IUnknown
	{
		static constexpr GUID iid() { return GUID{ 0x00000000, 0x0000, 0x0000, { 0xc0, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x46 } }; }
		virtual HRESULT WSPCALL QueryInterface( REFIID riid, void** ppvObject ) = 0;
		virtual uint32_t WSPCALL AddRef() = 0;
		virtual uint32_t WSPCALL Release() = 0;
	};
}

// The reference's iAudioReader hands out the Media Foundation source reader it wraps; its header only forward-declares the type
// (`struct IMFSourceReader;`, Whisper/API/iMediaFoundation.cl.h:5).  There is no Media Foundation here, so this library DEFINES that
// type for Linux clients: a pull source of 16 kHz mono f32 PCM in stream order (what Whisper/MF/PcmReader.cpp turns the MF samples
// into before anything else looks at them).
struct IMFSourceReader : public ComLight::IUnknown
{
	static constexpr GUID iid() { return GUID{ 0x6f1d2c3a, 0x52b7, 0x4e0c, { 0x9a, 0x41, 0x7b, 0x20, 0x0b, 0x2b, 0x20, 0x01 } }; }   // this library's own
	// Up to `capacity` samples into `mono`; *written == 0 with S_OK means the stream has ended.  Called from one thread at a time,
	// possibly not the thread that called runStreamed (MelStreamerThread reads ahead on a background thread).
	virtual HRESULT WSPCALL readPcm( float* mono, uint32_t capacity, uint32_t* written ) = 0;
};

namespace Whisper
{
	using whisper_token = int;
	struct iModel;
	struct iContext;

	// ---- TranscribeStructs.h ----
	struct sTimeSpan { uint64_t ticks; };   // 100 ns units
	struct sTimeInterval { sTimeSpan begin, end; };
	struct sSegment
	{
		const char* text;
		sTimeInterval time;
		uint32_t firstToken, countTokens;
	};
	enum eTokenFlags : uint32_t { None = 0, Special = 1 };
	struct sToken
	{
		const char* text;
		sTimeInterval time;
		float probability, probabilityTimestamp, ptsum, vlen;
		int id;
		eTokenFlags flags;
	};
	struct sTranscribeLength { uint32_t countSegments, countTokens; };
	enum struct eResultFlags : uint32_t { None = 0, Tokens = 1, Timestamps = 2, NewObject = 0x100 };
	enum struct eSpeakerChannel : uint8_t { Unsure = 0, Left = 1, Right = 2, NoStereoData = 0xFF };

	// ---- sFullParams.h ----
	enum struct eSamplingStrategy : int { Greedy, BeamSearch };
	using pfnNewSegment = HRESULT( * )( iContext* ctx, uint32_t n_new, void* user_data ) noexcept;
	using pfnEncoderBegin = HRESULT( * )( iContext* ctx, void* user_data ) noexcept;
	enum struct eFullParamsFlags : uint32_t
	{
		Translate = 1, NoContext = 2, SingleSegment = 4, PrintSpecial = 8, PrintProgress = 0x10, PrintRealtime = 0x20, PrintTimestamps = 0x40,
		TokenTimestamps = 0x100, SpeedupAudio = 0x200,
	};
	struct sFullParams
	{
		eSamplingStrategy strategy;
		int cpuThreads;
		int n_max_text_ctx;
		int offset_ms;
		int duration_ms;
		eFullParamsFlags flags;
		uint32_t language;
		float thold_pt, thold_ptsum;
		int max_len, max_tokens;
		struct { int n_past; } greedy;
		struct { int n_past, beam_width, n_best; } beam_search;
		int audio_ctx;
		const whisper_token* prompt_tokens;
		int prompt_n_tokens;
		pfnNewSegment new_segment_callback;
		void* new_segment_callback_user_data;
		pfnEncoderBegin encoder_begin_callback;
		void* encoder_begin_callback_user_data;
		bool flag( eFullParamsFlags f ) const { return 0 != ( (uint32_t)flags & (uint32_t)f ); }
	};
	inline uint32_t makeLanguageKey( const char* code )
	{
		uint32_t res = 0;
		for( uint32_t i = 0, shift = 0; i < 4 && code[ i ]; i++, shift += 8 ) res |= (uint32_t)(uint8_t)code[ i ] << shift;
		return res;
	}
	using pfnReportProgress = HRESULT( WSPCALL* )( double val, iContext* ctx, void* pv ) noexcept;
	struct sProgressSink { pfnReportProgress pfn; void* pv; };

	// ---- sModelSetup.h / sLoadModelCallbacks.h / loggerApi.h / SpecialTokens.h / sLanguageList.h ----
	enum struct eModelImplementation : uint32_t { GPU = 1, Hybrid = 2, Reference = 3, B200 = 4 };
	struct sModelSetup
	{
		eModelImplementation impl = eModelImplementation::GPU;
		uint32_t flags = 0;
		const wchar_t* adapter = nullptr;
	};
	using pfnListAdapters = void( WSPCALL* )( const wchar_t* name, void* pv );
	using pfnDecodedTokens = void( WSPCALL* )( const int* tokens, int tokensLength, void* pv );
	using pfnLoadProgress = HRESULT( WSPCALL* )( double val, void* pv ) noexcept;
	using pfnCancel = HRESULT( WSPCALL* )( void* pv ) noexcept;
	struct sLoadModelCallbacks { pfnLoadProgress progress; pfnCancel cancel; void* pv; };
	enum struct eLogLevel : uint8_t { Error = 0, Warning = 1, Info = 2, Debug = 3 };
	enum struct eLoggerFlags : uint8_t { UseStandardError = 1, SkipFormatMessage = 2 };
	using pfnLoggerSink = void( WSPCALL* )( void* context, eLogLevel lvl, const char* message );
	struct sLoggerSetup
	{
		pfnLoggerSink sink = nullptr;
		void* context = nullptr;
		eLogLevel level;
		eLoggerFlags flags = (eLoggerFlags)0;
	};
	struct SpecialTokens { int TranscriptionEnd, TranscriptionStart, PreviousWord, SentenceStart, Not, TranscriptionBegin, TaskTranslate, TaskTranscribe; };
	struct sLanguageEntry { uint32_t key; int id; const char* name; };
	struct sLanguageList { uint32_t length; const sLanguageEntry* pointer; };

	// ---- interfaces ----
	struct iAudioBuffer : public ComLight::IUnknown
	{
		static constexpr GUID iid() { return GUID{ 0x013583aa, 0xc9eb, 0x42bc, { 0x83, 0xdb, 0x63, 0x3c, 0x2c, 0x31, 0x70, 0x51 } }; }
		virtual uint32_t WSPCALL countSamples() const = 0;
		virtual const float* WSPCALL getPcmMono() const = 0;
		virtual const float* WSPCALL getPcmStereo() const = 0;
		virtual HRESULT WSPCALL getTime( int64_t& rdi ) const = 0;
	};
	struct iAudioReader : public ComLight::IUnknown
	{
		static constexpr GUID iid() { return GUID{ 0x35b988da, 0x04a6, 0x476a, { 0xa1, 0x93, 0xd8, 0x89, 0x1d, 0x5d, 0xc3, 0x90 } }; }
		virtual HRESULT WSPCALL getDuration( int64_t& rdi ) const = 0;          // 100 ns ticks; the stream is floor( ticks / 100000 ) mel frames long (PcmReader.cpp:247-272)
		virtual HRESULT WSPCALL getReader( IMFSourceReader** pp ) const = 0;    // AddRef'ed
		virtual HRESULT WSPCALL requestedStereo() const = 0;                    // S_OK / S_FALSE; stereo (diarisation) is not delivered here
	};
	// ---- MfStructs.h: live capture ----
	enum struct eCaptureFlags : uint32_t { Stereo = 1 };
	struct sCaptureParams
	{
		float minDuration = 2.0f;
		float maxDuration = 3.0f;
		float dropStartSilence = 0.25f;
		float pauseDuration = 0.333f;
		uint32_t flags = 0;
	};
	enum struct eCaptureStatus : uint8_t { Listening = 1, Voice = 2, Transcribing = 4, Stalled = 0x80 };
	using pfnShouldCancel = HRESULT( WSPCALL* )( void* pv ) noexcept;                            // S_OK to continue, S_FALSE to stop the capture session
	using pfnCaptureStatus = HRESULT( WSPCALL* )( void* pv, eCaptureStatus status ) noexcept;
	struct sCaptureCallbacks
	{
		pfnShouldCancel shouldCancel;
		pfnCaptureStatus captureStatus;
		void* pv;
	};
	struct iAudioCapture : public ComLight::IUnknown
	{
		static constexpr GUID iid() { return GUID{ 0x747752c2, 0xd9fd, 0x40df, { 0x88, 0x47, 0x58, 0x3c, 0x78, 0x1b, 0xf0, 0x13 } }; }
		virtual HRESULT WSPCALL getReader( IMFSourceReader** pp ) const = 0;
		virtual const sCaptureParams& WSPCALL getParams() const = 0;
	};
	struct sCaptureDevice { const wchar_t* displayName; const wchar_t* endpoint; };                 // MfStructs.h:5-14
	using pfnFoundCaptureDevices = HRESULT( WSPCALL* )( int len, const sCaptureDevice* buffer, void* pv );
#ifndef _MSC_VER
	using LPCTSTR = const char*;                                                                     // ComLightLib/comLightCommon.h:8
#endif
	// The reference's media layer (Whisper/API/iMediaFoundation.cl.h:36-49).  The Linux object behind initMediaFoundation decodes
	// RIFF/WAVE (16-bit PCM or 32-bit float, any rate and channel count) instead of everything Media Foundation can open; it has no
	// capture devices (listCaptureDevices reports none, openCaptureDevice answers E_NOTIMPL — use createAudioCapture).
	struct iMediaFoundation : public ComLight::IUnknown
	{
		static constexpr GUID iid() { return GUID{ 0xfb9763a5, 0xd77d, 0x4b6e, { 0xaf, 0xf8, 0xf4, 0x94, 0x81, 0x3c, 0xeb, 0xd8 } }; }
		virtual HRESULT WSPCALL loadAudioFile( LPCTSTR path, bool stereo, iAudioBuffer** pp ) const = 0;
		virtual HRESULT WSPCALL openAudioFile( LPCTSTR path, bool stereo, iAudioReader** pp ) = 0;
		virtual HRESULT WSPCALL loadAudioFileData( const void* data, uint64_t size, bool stereo, iAudioReader** pp ) = 0;
		virtual HRESULT WSPCALL listCaptureDevices( pfnFoundCaptureDevices pfn, void* pv ) = 0;
		virtual HRESULT WSPCALL openCaptureDevice( LPCTSTR endpoint, const sCaptureParams& captureParams, iAudioCapture** pp ) = 0;
	};

	struct iTranscribeResult : public ComLight::IUnknown
	{
		static constexpr GUID iid() { return GUID{ 0x2871a73f, 0x5ce3, 0x48f8, { 0x87, 0x79, 0x65, 0x82, 0xee, 0x11, 0x93, 0x5e } }; }
		virtual HRESULT WSPCALL getSize( sTranscribeLength& rdi ) const = 0;
		virtual const sSegment* WSPCALL getSegments() const = 0;
		virtual const sToken* WSPCALL getTokens() const = 0;
	};

	struct iContext : public ComLight::IUnknown
	{
		static constexpr GUID iid() { return GUID{ 0xb9956374, 0x3b18, 0x4943, { 0x90, 0xf2, 0x2a, 0xb1, 0x8a, 0x40, 0x45, 0x37 } }; }
		virtual HRESULT WSPCALL runFull( const sFullParams& params, const iAudioBuffer* buffer ) = 0;
		virtual HRESULT WSPCALL runStreamed( const sFullParams& params, const sProgressSink& progress, const iAudioReader* reader ) = 0;
		virtual HRESULT WSPCALL runCapture( const sFullParams& params, const sCaptureCallbacks& callbacks, const iAudioCapture* reader ) = 0;
		virtual HRESULT WSPCALL getResults( eResultFlags flags, iTranscribeResult** pp ) const = 0;
		virtual HRESULT WSPCALL detectSpeaker( const sTimeInterval& time, eSpeakerChannel& result ) const = 0;
		virtual HRESULT WSPCALL getModel( iModel** pp ) = 0;
		virtual HRESULT WSPCALL fullDefaultParams( eSamplingStrategy strategy, sFullParams* rdi ) = 0;
		virtual HRESULT WSPCALL timingsPrint() = 0;
		virtual HRESULT WSPCALL timingsReset() = 0;
	};

	struct iModel : public ComLight::IUnknown
	{
		static constexpr GUID iid() { return GUID{ 0xabefb4c9, 0xe8d8, 0x46a3, { 0x87, 0x47, 0x5a, 0xfb, 0xad, 0xef, 0x1a, 0xdb } }; }
		virtual HRESULT WSPCALL createContext( iContext** pp ) = 0;
		virtual HRESULT WSPCALL tokenize( const char* text, pfnDecodedTokens pfn, void* pv ) = 0;
		virtual HRESULT WSPCALL isMultilingual() = 0;
		virtual HRESULT WSPCALL getSpecialTokens( SpecialTokens& rdi ) = 0;
		virtual const char* WSPCALL stringFromToken( whisper_token token ) = 0;
		virtual HRESULT WSPCALL clone( iModel** rdi ) = 0;
	};

	// ---- exported functions (Whisper/whisper.def) ----
	extern "C++" {
	HRESULT WSPCALL setupLogger( const sLoggerSetup& setup );
	HRESULT WSPCALL loadModel( const wchar_t* path, const sModelSetup& setup, const sLoadModelCallbacks* callbacks, iModel** pp );
	uint32_t WSPCALL findLanguageKeyW( const wchar_t* lang );
	uint32_t WSPCALL findLanguageKeyA( const char* lang );
	HRESULT WSPCALL getSupportedLanguages( sLanguageList& rdi );
	HRESULT WSPCALL listGPUs( pfnListAdapters pfn, void* pv );
	HRESULT WSPCALL initMediaFoundation( iMediaFoundation** pp );   // the WAV-only media layer described at iMediaFoundation above
	// not in the reference: an iAudioBuffer over caller-owned 16 kHz mono f32 PCM (replaces the Media Foundation loader, Whisper/MF/)
	HRESULT WSPCALL createAudioBuffer( const float* pcmMono, uint32_t countSamples, iAudioBuffer** pp );
	// the same with the interleaved left/right samples kept next to the mono mix (loadAudioFile( path, stereo = true ), the --diarize
	// option of the reference CLI): what iContext::detectSpeaker compares
	HRESULT WSPCALL createAudioBufferStereo( const float* pcmMono, const float* pcmStereo, uint32_t countSamples, iAudioBuffer** pp );
	// not in the reference either: an iAudioReader over a caller-supplied pull callback (replaces iMediaFoundation::openAudioFile /
	// loadAudioFileData, Whisper/API/iMediaFoundation.cl.h:38-39).  `durationTicks` announces the stream length (100 ns units).
	using pfnReadPcm = HRESULT( WSPCALL* )( float* mono, uint32_t capacity, uint32_t* written, void* pv ) noexcept;
	HRESULT WSPCALL createAudioReader( pfnReadPcm pfn, void* pv, int64_t durationTicks, iAudioReader** pp );
	// ... and an iAudioCapture over the same kind of callback, for a live source (replaces iMediaFoundation::openCaptureDevice, :42).
	// The callback blocks until audio is available; returning *written == 0 ends the session with E_EOF like a device that went away.
	HRESULT WSPCALL createAudioCapture( pfnReadPcm pfn, void* pv, const sCaptureParams& captureParams, iAudioCapture** pp );
	}
}
