// TEST INFRASTRUCTURE — flat C helpers over the COM-style surface (include/whisper_b200_com.h) so that pytest can drive C++ vtables
// through ctypes.  Built into tests/boundary/_build/libwspc_test.so and linked against libwhisper_b200.so: it uses nothing but the
// public header, exactly like a client application would (Examples/main/main.cpp:210-318 of the reference).
#include "whisper_b200_com.h"
#include "../../whisper_b200/csrc/capture_loop.h"   // host-only header: used to predict where the capture loop cuts (wspc_capture_cuts)
#include <atomic>
#include <chrono>
#include <string>
#include <thread>
#include <vector>
using namespace Whisper;

// every call goes loadModel -> createContext -> fullDefaultParams -> runFull -> getResults exactly as the reference's CLI does
namespace
{
	struct Session
	{
		iModel* model = nullptr;
		iContext* context = nullptr;
		iTranscribeResult* result = nullptr;
		std::vector<int> segCallbackCounts;
		int maxLen = 0;
		// runCapture: segments collected from inside new_segment_callback (on the library's transcriber thread)
		struct CapturedSeg { uint64_t t0, t1; std::string text; std::vector<int> tokens; };
		std::vector<CapturedSeg> captured;
	};
	HRESULT captureSegCallback( iContext* ctx, uint32_t nNew, void* pv ) noexcept
	{
		Session* s = static_cast<Session*>( pv );
		iTranscribeResult* res = nullptr;
		if( FAILED( ctx->getResults( (eResultFlags)( (uint32_t)eResultFlags::Tokens | (uint32_t)eResultFlags::Timestamps ), &res ) ) ) return E_FAIL;
		sTranscribeLength len;
		res->getSize( len );
		for( uint32_t i = len.countSegments - nNew; i < len.countSegments; i++ )
		{
			const sSegment& seg = res->getSegments()[ i ];
			Session::CapturedSeg c{ seg.time.begin.ticks, seg.time.end.ticks, seg.text, {} };
			for( uint32_t j = 0; j < seg.countTokens; j++ ) c.tokens.push_back( res->getTokens()[ seg.firstToken + j ].id );
			s->captured.push_back( std::move( c ) );
		}
		res->Release();
		return S_OK;
	}
	// a live source played from an array: it delivers ragged blocks and — like a real-time device feeding a transcriber that is faster
	// than real time — never runs ahead of a transcription in flight, which makes the utterance cuts deterministic
	struct LiveSource
	{
		const float* pcm; uint32_t n, pos = 0, lcg = 4242u;
		std::atomic<uint8_t> status{ 0 };
		int transcriptions = 0, stalls = 0;
		LiveSource( const float* p, uint32_t count ) : pcm( p ), n( count ) {}
		HRESULT read( float* mono, uint32_t capacity, uint32_t* written )
		{
			for( int waited = 0; ( status.load() & (uint8_t)eCaptureStatus::Transcribing ) && waited < 300000; waited++ )   // at most ~30 s: a test must not hang
				std::this_thread::sleep_for( std::chrono::microseconds( 100 ) );
			lcg = lcg * 1664525u + 1013904223u;
			uint32_t want = 100 + ( lcg >> 8 ) % 900;
			if( want > capacity ) want = capacity;
			if( want > n - pos ) want = n - pos;
			if( want ) memcpy( mono, pcm + pos, (size_t)want * 4 );
			pos += want;
			*written = want;
			return S_OK;
		}
		HRESULT onStatus( uint8_t bits )
		{
			const uint8_t old = status.exchange( bits );
			if( ( bits & (uint8_t)eCaptureStatus::Transcribing ) && !( old & (uint8_t)eCaptureStatus::Transcribing ) ) transcriptions++;
			if( ( bits & (uint8_t)eCaptureStatus::Stalled ) && !( old & (uint8_t)eCaptureStatus::Stalled ) ) stalls++;
			return S_OK;
		}
		bool finished() const { return pos >= n && !( status.load() & (uint8_t)eCaptureStatus::Transcribing ); }
	};
	HRESULT segCallback( iContext*, uint32_t nNew, void* pv ) noexcept
	{
		static_cast<Session*>( pv )->segCallbackCounts.push_back( (int)nNew );
		return S_OK;
	}
}

extern "C" {

int32_t wspc_open( const char* modelPathUtf8, int32_t device, void** out )
{
	if( !modelPathUtf8 || !out ) return E_POINTER;
	std::wstring w;
	for( const char* p = modelPathUtf8; *p; p++ ) w.push_back( (wchar_t)(unsigned char)*p );   // test paths are ASCII
	std::wstring adapter = std::to_wstring( device );
	sModelSetup setup;
	setup.impl = eModelImplementation::B200;
	setup.adapter = adapter.c_str();
	Session* s = new Session();
	HRESULT hr = loadModel( w.c_str(), setup, nullptr, &s->model );
	if( SUCCEEDED( hr ) ) hr = s->model->createContext( &s->context );
	if( FAILED( hr ) )
	{
		if( s->model ) s->model->Release();
		delete s;
		return hr;
	}
	*out = s;
	return S_OK;
}
void wspc_close( void* h )
{
	Session* s = static_cast<Session*>( h );
	if( !s ) return;
	if( s->result ) s->result->Release();
	if( s->context ) s->context->Release();
	if( s->model ) s->model->Release();
	delete s;
}
// flags = eFullParamsFlags bits; language = code such as "en" or "auto"
int32_t wspc_run_full( void* h, const float* pcm, int32_t nSamples, uint32_t flags, const char* language, int32_t maxTokens, int32_t cpuThreads,
	int32_t offsetMs, int32_t durationMs, const int32_t* promptTokens, int32_t nPromptTokens )
{
	Session* s = static_cast<Session*>( h );
	if( !s ) return E_POINTER;
	sFullParams p;
	HRESULT hr = s->context->fullDefaultParams( eSamplingStrategy::Greedy, &p );
	if( FAILED( hr ) ) return hr;
	p.flags = (eFullParamsFlags)flags;
	p.language = ( language && strcmp( language, "auto" ) != 0 ) ? findLanguageKeyA( language ) : makeLanguageKey( "auto" );
	p.max_tokens = maxTokens;
	p.max_len = s->maxLen;
	p.cpuThreads = cpuThreads;
	p.offset_ms = offsetMs;
	p.duration_ms = durationMs;
	p.prompt_tokens = promptTokens;
	p.prompt_n_tokens = nPromptTokens;
	p.new_segment_callback = &segCallback;
	p.new_segment_callback_user_data = s;
	s->segCallbackCounts.clear();
	iAudioBuffer* buf = nullptr;
	hr = createAudioBuffer( pcm, (uint32_t)nSamples, &buf );
	if( FAILED( hr ) ) return hr;
	hr = s->context->runFull( p, buf );
	buf->Release();
	if( FAILED( hr ) ) return hr;
	if( s->result ) { s->result->Release(); s->result = nullptr; }
	const HRESULT hr2 = s->context->getResults( (eResultFlags)( (uint32_t)eResultFlags::Tokens | (uint32_t)eResultFlags::Timestamps ), &s->result );
	return FAILED( hr2 ) ? hr2 : hr;
}
// iContext::runStreamed over an iAudioReader made with createAudioReader: the PCM is handed out in ragged blocks (1..maxBlock samples,
// never more than asked for), the way a decoder delivers it.  progressOut[0] = number of progress calls, [1] = 1 if the per-window
// values never decreased (the last window may overshoot 1.0 when the final seek passes the end, as in the reference's formula) and
// the closing call reported exactly 1.0, [2] = number of read calls.
int32_t wspc_run_streamed( void* h, const float* pcm, int32_t nSamples, uint32_t flags, const char* language, int32_t maxTokens, int32_t cpuThreads,
	int32_t offsetMs, int32_t durationMs, int32_t maxBlock, int32_t* progressOut )
{
	Session* s = static_cast<Session*>( h );
	if( !s ) return E_POINTER;
	sFullParams p;
	HRESULT hr = s->context->fullDefaultParams( eSamplingStrategy::Greedy, &p );
	if( FAILED( hr ) ) return hr;
	p.flags = (eFullParamsFlags)flags;
	p.language = ( language && strcmp( language, "auto" ) != 0 ) ? findLanguageKeyA( language ) : makeLanguageKey( "auto" );
	p.max_tokens = maxTokens;
	p.cpuThreads = cpuThreads;
	p.offset_ms = offsetMs;
	p.duration_ms = durationMs;
	p.new_segment_callback = &segCallback;
	p.new_segment_callback_user_data = s;
	s->segCallbackCounts.clear();
	struct Source { const float* pcm; uint32_t n, pos, maxBlock, calls, lcg; } src{ pcm, (uint32_t)nSamples, 0, (uint32_t)( maxBlock > 0 ? maxBlock : 1 ), 0, 12345u };
	auto readFn = []( float* mono, uint32_t capacity, uint32_t* written, void* pv ) noexcept -> HRESULT {
		Source* k = static_cast<Source*>( pv );
		k->calls++;
		k->lcg = k->lcg * 1664525u + 1013904223u;
		uint32_t want = 1 + ( k->lcg >> 8 ) % k->maxBlock;
		if( want > capacity ) want = capacity;
		if( want > k->n - k->pos ) want = k->n - k->pos;
		if( want ) memcpy( mono, k->pcm + k->pos, (size_t)want * 4 );
		k->pos += want;
		*written = want;
		return S_OK;
	};
	struct Progress { int calls = 0; double last = -1.0, prev = -1.0; bool monotonic = true; } prog;
	sProgressSink sink;
	sink.pfn = []( double val, iContext*, void* pv ) noexcept -> HRESULT {
		Progress* g = static_cast<Progress*>( pv );
		if( g->last < g->prev ) g->monotonic = false;   // judged one call late: the closing 1.0 is exempt
		g->prev = g->last;
		g->last = val;
		g->calls++;
		return S_OK;
	};
	sink.pv = &prog;
	iAudioReader* reader = nullptr;
	hr = createAudioReader( readFn, &src, (int64_t)( (uint32_t)nSamples / 160 ) * 100000, &reader );
	if( FAILED( hr ) ) return hr;
	hr = s->context->runStreamed( p, sink, reader );
	reader->Release();
	if( progressOut )
	{
		progressOut[ 0 ] = prog.calls;
		progressOut[ 1 ] = ( prog.monotonic && prog.last == 1.0 ) ? 1 : 0;
		progressOut[ 2 ] = (int32_t)src.calls;
	}
	if( FAILED( hr ) ) return hr;
	if( s->result ) { s->result->Release(); s->result = nullptr; }
	const HRESULT hr2 = s->context->getResults( (eResultFlags)( (uint32_t)eResultFlags::Tokens | (uint32_t)eResultFlags::Timestamps ), &s->result );
	return FAILED( hr2 ) ? hr2 : hr;
}
// iContext::runCapture over an iAudioCapture made with createAudioCapture.  info[0] = segments collected, [1] = transcriptions started,
// [2] = stalls, [3] = samples the source handed out.
int32_t wspc_run_capture( void* h, const float* pcm, int32_t nSamples, uint32_t flags, const char* language, int32_t cpuThreads, float minDuration, float maxDuration, int32_t* info )
{
	Session* s = static_cast<Session*>( h );
	if( !s ) return E_POINTER;
	sFullParams p;
	HRESULT hr = s->context->fullDefaultParams( eSamplingStrategy::Greedy, &p );
	if( FAILED( hr ) ) return hr;
	p.flags = (eFullParamsFlags)flags;
	p.language = findLanguageKeyA( language );
	p.cpuThreads = cpuThreads;
	p.new_segment_callback = &captureSegCallback;
	p.new_segment_callback_user_data = s;
	s->captured.clear();
	LiveSource src( pcm, (uint32_t)nSamples );
	sCaptureParams cp;
	cp.minDuration = minDuration;
	cp.maxDuration = maxDuration;
	iAudioCapture* cap = nullptr;
	hr = createAudioCapture( []( float* m, uint32_t c, uint32_t* w, void* pv ) noexcept -> HRESULT { return static_cast<LiveSource*>( pv )->read( m, c, w ); }, &src, cp, &cap );
	if( FAILED( hr ) ) return hr;
	sCaptureCallbacks cb;
	cb.pv = &src;
	cb.shouldCancel = []( void* pv ) noexcept -> HRESULT { return static_cast<LiveSource*>( pv )->finished() ? S_FALSE : S_OK; };
	cb.captureStatus = []( void* pv, eCaptureStatus st ) noexcept -> HRESULT { return static_cast<LiveSource*>( pv )->onStatus( (uint8_t)st ); };
	hr = s->context->runCapture( p, cb, cap );
	cap->Release();
	if( info )
	{
		info[ 0 ] = (int32_t)s->captured.size(); info[ 1 ] = src.transcriptions; info[ 2 ] = src.stalls; info[ 3 ] = (int32_t)src.pos;
	}
	return hr;
}
int64_t wspc_captured_t0( void* h, int32_t i ) { return (int64_t) static_cast<Session*>( h )->captured[ i ].t0; }   // 100 ns ticks
int64_t wspc_captured_t1( void* h, int32_t i ) { return (int64_t) static_cast<Session*>( h )->captured[ i ].t1; }
const char* wspc_captured_text( void* h, int32_t i ) { return static_cast<Session*>( h )->captured[ i ].text.c_str(); }
int32_t wspc_captured_n_tokens( void* h, int32_t i ) { return (int32_t) static_cast<Session*>( h )->captured[ i ].tokens.size(); }
int32_t wspc_captured_token( void* h, int32_t i, int32_t j ) { return static_cast<Session*>( h )->captured[ i ].tokens[ j ]; }
// Where the capture loop cuts the same signal, predicted on the CPU with the same (host-only) loop and a transcriber that does nothing:
// starts[i], sizes[i] in samples; returns the number of utterances.
int32_t wspc_capture_cuts( const float* pcm, int32_t nSamples, float minDuration, float maxDuration, int64_t* starts, int32_t* sizes, int32_t cap )
{
	LiveSource src( pcm, (uint32_t)nSamples );
	sCaptureParams cp;
	int32_t n = 0;
	{
		wsp::CaptureLoop loop(
			[ &src ]( float* d, uint32_t c, uint32_t* w ) -> int32_t { return src.read( d, c, w ); },
			[ &src ]( uint8_t bits ) -> int32_t { return src.onStatus( bits ); },
			[ & ]( const std::vector<float>& u, int64_t first ) -> int32_t { if( n < cap ) { starts[ n ] = first; sizes[ n ] = (int32_t)u.size(); } n++; return 0; },
			wsp::CaptureLoop::settingsFromSeconds( minDuration, maxDuration, cp.dropStartSilence, cp.pauseDuration ) );
		loop.startup();
		while( !src.finished() && loop.step() >= 0 ) {}
	}
	return n;
}
// runFull on a buffer that keeps its interleaved stereo samples; iContext::detectSpeaker is asked about every new segment from inside
// new_segment_callback (the only place it works, like in the reference).  speakers[i] = eSpeakerChannel of segment i; returns the
// HRESULT of runFull, or of the first failing detectSpeaker.  pcmStereo == NULL uses a mono-only buffer.  *afterRun receives what
// detectSpeaker answers once the run is over.
int32_t wspc_run_full_stereo( void* h, const float* pcmMono, const float* pcmStereo, int32_t nSamples, uint32_t flags, int32_t durationMs, int32_t* speakers, int32_t cap, int32_t* afterRun )
{
	Session* s = static_cast<Session*>( h );
	if( !s ) return E_POINTER;
	sFullParams p;
	HRESULT hr = s->context->fullDefaultParams( eSamplingStrategy::Greedy, &p );
	if( FAILED( hr ) ) return hr;
	p.flags = (eFullParamsFlags)flags;
	p.duration_ms = durationMs;
	struct Log { int32_t* out; int32_t cap, n; HRESULT failed; } log{ speakers, cap, 0, S_OK };
	p.new_segment_callback = []( iContext* ctx, uint32_t nNew, void* pv ) noexcept -> HRESULT {
		Log* g = static_cast<Log*>( pv );
		iTranscribeResult* res = nullptr;
		if( FAILED( ctx->getResults( eResultFlags::Timestamps, &res ) ) ) return E_FAIL;
		sTranscribeLength len;
		res->getSize( len );
		for( uint32_t i = len.countSegments - nNew; i < len.countSegments; i++ )
		{
			eSpeakerChannel ch = eSpeakerChannel::Unsure;
			const HRESULT hr = ctx->detectSpeaker( res->getSegments()[ i ].time, ch );
			if( FAILED( hr ) ) g->failed = hr;
			if( g->n < g->cap ) g->out[ g->n ] = (int32_t)(uint8_t)ch;
			g->n++;
		}
		res->Release();
		return S_OK;
	};
	p.new_segment_callback_user_data = &log;
	iAudioBuffer* buf = nullptr;
	hr = pcmStereo ? createAudioBufferStereo( pcmMono, pcmStereo, (uint32_t)nSamples, &buf ) : createAudioBuffer( pcmMono, (uint32_t)nSamples, &buf );
	if( FAILED( hr ) ) return hr;
	hr = s->context->runFull( p, buf );
	buf->Release();
	if( afterRun )
	{
		sTimeInterval t{ { 0 }, { 10000000 } };
		eSpeakerChannel ch;
		*afterRun = s->context->detectSpeaker( t, ch );
	}
	if( FAILED( hr ) ) return hr;
	if( FAILED( log.failed ) ) return log.failed;
	if( s->result ) { s->result->Release(); s->result = nullptr; }
	const HRESULT hr2 = s->context->getResults( (eResultFlags)( (uint32_t)eResultFlags::Tokens | (uint32_t)eResultFlags::Timestamps ), &s->result );
	return FAILED( hr2 ) ? hr2 : log.n;
}
// ---- the media layer: initMediaFoundation -> loadAudioFile / openAudioFile / loadAudioFileData (no GPU involved) ----
// mode 0: loadAudioFile (whole file -> iAudioBuffer); 1: openAudioFile, 2: loadAudioFileData on the file's bytes — both read to the end
// through iAudioReader::getReader()->readPcm in blocks of `block` samples.  mono[cap] receives the samples, stereo[2*cap] (nullable) the
// pairs (mode 0 only); info[0] = 1 if the object reports stereo, info[1..2] = getDuration ticks (low / high word).  Returns the sample
// count or a negative HRESULT.
int32_t wspc_media_decode( const char* path, int32_t mode, int32_t wantStereo, int32_t block, float* mono, float* stereo, int32_t cap, int32_t* info )
{
	iMediaFoundation* mf = nullptr;
	HRESULT hr = initMediaFoundation( &mf );
	if( FAILED( hr ) ) return hr;
	int32_t n = 0;
	if( mode == 0 )
	{
		iAudioBuffer* buf = nullptr;
		hr = mf->loadAudioFile( path, wantStereo != 0, &buf );
		if( SUCCEEDED( hr ) )
		{
			n = (int32_t)buf->countSamples();
			if( n > cap ) n = cap;
			memcpy( mono, buf->getPcmMono(), (size_t)n * 4 );
			const float* st = buf->getPcmStereo();
			if( info ) info[ 0 ] = st ? 1 : 0;
			if( st && stereo ) memcpy( stereo, st, (size_t)n * 8 );
			buf->Release();
		}
	}
	else
	{
		iAudioReader* reader = nullptr;
		std::vector<uint8_t> bytes;
		if( mode == 1 ) hr = mf->openAudioFile( path, wantStereo != 0, &reader );
		else
		{
			FILE* f = fopen( path, "rb" );
			if( !f ) { mf->Release(); return E_INVALIDARG; }
			fseek( f, 0, SEEK_END ); const long sz = ftell( f ); fseek( f, 0, SEEK_SET );
			bytes.resize( (size_t)sz );
			const bool ok = fread( bytes.data(), 1, bytes.size(), f ) == bytes.size();
			fclose( f );
			if( !ok ) { mf->Release(); return E_FAIL; }
			hr = mf->loadAudioFileData( bytes.data(), bytes.size(), wantStereo != 0, &reader );
			bytes.assign( bytes.size(), 0xAB );   // the library must have taken its own copy
		}
		if( SUCCEEDED( hr ) )
		{
			int64_t ticks = 0;
			reader->getDuration( ticks );
			if( info ) { info[ 0 ] = reader->requestedStereo() == S_OK ? 1 : 0; info[ 1 ] = (int32_t)( ticks & 0xFFFFFFFF ); info[ 2 ] = (int32_t)( ticks >> 32 ); }
			IMFSourceReader* src = nullptr;
			hr = reader->getReader( &src );
			while( SUCCEEDED( hr ) && n < cap )
			{
				uint32_t got = 0;
				const uint32_t want = (uint32_t)( cap - n < block ? cap - n : block );
				hr = src->readPcm( mono + n, want, &got );
				if( FAILED( hr ) || got == 0 ) break;
				n += (int32_t)got;
			}
			if( src ) src->Release();
			reader->Release();
		}
	}
	int devices = -1;
	mf->listCaptureDevices( []( int len, const sCaptureDevice*, void* pv ) -> HRESULT { *static_cast<int*>( pv ) = len; return S_OK; }, &devices );
	iAudioCapture* cap0 = nullptr;
	const HRESULT hrCap = mf->openCaptureDevice( "default", sCaptureParams{}, &cap0 );
	mf->Release();
	if( devices != 0 || hrCap != E_NOTIMPL ) return E_UNEXPECTED;
	return FAILED( hr ) ? hr : n;
}
void wspc_set_max_len( void* h, int32_t maxLen ) { static_cast<Session*>( h )->maxLen = maxLen; }
int64_t wspc_token_t0( void* h, int32_t i, int32_t j )
{
	Session* s = static_cast<Session*>( h );
	return (int64_t)s->result->getTokens()[ s->result->getSegments()[ i ].firstToken + j ].time.begin.ticks / 100000;
}
int64_t wspc_token_t1( void* h, int32_t i, int32_t j )
{
	Session* s = static_cast<Session*>( h );
	return (int64_t)s->result->getTokens()[ s->result->getSegments()[ i ].firstToken + j ].time.end.ticks / 100000;
}
int32_t wspc_segment_callback_total( void* h )
{
	Session* s = static_cast<Session*>( h );
	int n = 0;
	for( int v : s->segCallbackCounts ) n += v;
	return n;
}
int32_t wspc_n_segments( void* h )
{
	Session* s = static_cast<Session*>( h );
	if( !s || !s->result ) return 0;
	sTranscribeLength len;
	s->result->getSize( len );
	return (int32_t)len.countSegments;
}
int32_t wspc_n_segment_callbacks( void* h ) { Session* s = static_cast<Session*>( h ); return s ? (int32_t)s->segCallbackCounts.size() : 0; }
int64_t wspc_segment_t0( void* h, int32_t i ) { return (int64_t)( static_cast<Session*>( h )->result->getSegments()[ i ].time.begin.ticks / 100000 ); }
int64_t wspc_segment_t1( void* h, int32_t i ) { return (int64_t)( static_cast<Session*>( h )->result->getSegments()[ i ].time.end.ticks / 100000 ); }
const char* wspc_segment_text( void* h, int32_t i ) { return static_cast<Session*>( h )->result->getSegments()[ i ].text; }
int32_t wspc_segment_n_tokens( void* h, int32_t i ) { return (int32_t) static_cast<Session*>( h )->result->getSegments()[ i ].countTokens; }
int32_t wspc_token_id( void* h, int32_t i, int32_t j )
{
	Session* s = static_cast<Session*>( h );
	const sSegment& seg = s->result->getSegments()[ i ];
	return s->result->getTokens()[ seg.firstToken + j ].id;
}
float wspc_token_p( void* h, int32_t i, int32_t j )
{
	Session* s = static_cast<Session*>( h );
	const sSegment& seg = s->result->getSegments()[ i ];
	return s->result->getTokens()[ seg.firstToken + j ].probability;
}
int32_t wspc_token_flags( void* h, int32_t i, int32_t j )
{
	Session* s = static_cast<Session*>( h );
	const sSegment& seg = s->result->getSegments()[ i ];
	return (int32_t)s->result->getTokens()[ seg.firstToken + j ].flags;
}
// iModel surface
int32_t wspc_tokenize( void* h, const char* text, int32_t* dst, int32_t cap )
{
	Session* s = static_cast<Session*>( h );
	struct Sink { int32_t* dst; int32_t cap; int32_t n; } sink{ dst, cap, 0 };
	auto cb = []( const int* tokens, int n, void* pv ) {
		Sink* k = static_cast<Sink*>( pv );
		for( int i = 0; i < n && k->n < k->cap; i++ ) k->dst[ k->n++ ] = tokens[ i ];
	};
	const HRESULT hr = s->model->tokenize( text, cb, &sink );
	return FAILED( hr ) ? hr : sink.n;
}
const char* wspc_string_from_token( void* h, int32_t id ) { return static_cast<Session*>( h )->model->stringFromToken( id ); }
int32_t wspc_is_multilingual( void* h ) { return static_cast<Session*>( h )->model->isMultilingual() == S_OK ? 1 : 0; }
int32_t wspc_special_tokens( void* h, int32_t* out8 )
{
	SpecialTokens st;
	const HRESULT hr = static_cast<Session*>( h )->model->getSpecialTokens( st );
	memcpy( out8, &st, sizeof( st ) );
	return hr;
}
int32_t wspc_query_interfaces( void* h )
{
	// IUnknown plumbing: QueryInterface round trips and reference counts behave like COM
	Session* s = static_cast<Session*>( h );
	void* p = nullptr;
	if( s->model->QueryInterface( iModel::iid(), &p ) != S_OK || p != s->model ) return 1;
	s->model->Release();
	if( s->model->QueryInterface( ComLight::IUnknown::iid(), &p ) != S_OK ) return 2;
	s->model->Release();
	if( s->model->QueryInterface( iContext::iid(), &p ) != E_NOINTERFACE || p != nullptr ) return 3;
	iModel* m2 = nullptr;
	if( s->context->getModel( &m2 ) != S_OK || m2 != s->model ) return 4;
	m2->Release();
	iModel* clone = nullptr;
	if( s->model->clone( &clone ) != S_OK || !clone ) return 5;
	clone->Release();
	sProgressSink sink{ nullptr, nullptr };
	sFullParams fp;
	s->context->fullDefaultParams( eSamplingStrategy::Greedy, &fp );
	if( s->context->runStreamed( fp, sink, nullptr ) != E_POINTER ) return 6;
	return 0;
}
uint32_t wspc_find_language_key( const char* lang ) { return findLanguageKeyA( lang ); }
int32_t wspc_language_count( void )
{
	sLanguageList l;
	getSupportedLanguages( l );
	return (int32_t)l.length;
}

} // extern "C"
