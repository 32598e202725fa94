// TEST INFRASTRUCTURE.  Prints the ABI-relevant facts of the COM-style surface — struct sizes, field offsets, enum values, interface GUIDs
// and vtable slot numbers — for ONE of the two declarations of it:
//   -DUSE_REFERENCE_HEADERS -I<reference root> -D__stdcall= -D__cdecl= : the reference's own headers (Whisper/API/*.h + ComLightLib/)
//   otherwise                                                          : include/whisper_b200_com.h
// tests/test_boundary.py compiles it both ways and requires identical output.
#include <stddef.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>   // the reference headers use strlen without including it (MSVC pulls it in transitively)
#ifdef USE_REFERENCE_HEADERS
#include "Whisper/API/whisperComLight.h"
#include "Whisper/API/sFullParams.h"
#else
#include "whisper_b200_com.h"
#endif
using namespace Whisper;

#define SZ( T ) printf( "sizeof(%s) = %zu align %zu\n", #T, sizeof( T ), alignof( T ) )
#define OFF( T, f ) printf( "  %s.%s @ %zu size %zu\n", #T, #f, offsetof( T, f ), sizeof( ( (T*)0 )->f ) )
#define ENUMV( e ) printf( "  %s = %lld\n", #e, (long long)( e ) )

static void guid( const char* name, const GUID& g )
{
	const unsigned char* p = (const unsigned char*)&g;
	printf( "iid(%s) =", name );
	for( size_t i = 0; i < sizeof( GUID ); i++ ) printf( " %02x", p[ i ] );
	printf( "\n" );
}

// vtable slot of a virtual method = index of the first difference between two fake objects' call logs is overkill; the Itanium ABI
// encodes a pointer-to-virtual-member-function as 1 + slot * sizeof(void*)
template<class PMF>
static long slot( PMF pmf )
{
	union { PMF p; struct { intptr_t ptr; intptr_t adj; } raw; } u;
	u.raw.ptr = 0; u.raw.adj = 0;
	u.p = pmf;
	return (long)( ( u.raw.ptr - 1 ) / (intptr_t)sizeof( void* ) );
}
#define SLOT( I, m ) printf( "  slot %s::%s = %ld\n", #I, #m, slot( &I::m ) )

int main()
{
	SZ( sFullParams );
	OFF( sFullParams, strategy ); OFF( sFullParams, cpuThreads ); OFF( sFullParams, n_max_text_ctx ); OFF( sFullParams, offset_ms ); OFF( sFullParams, duration_ms );
	OFF( sFullParams, flags ); OFF( sFullParams, language ); OFF( sFullParams, thold_pt ); OFF( sFullParams, thold_ptsum ); OFF( sFullParams, max_len );
	OFF( sFullParams, max_tokens ); OFF( sFullParams, greedy ); OFF( sFullParams, beam_search ); OFF( sFullParams, audio_ctx ); OFF( sFullParams, prompt_tokens );
	OFF( sFullParams, prompt_n_tokens ); OFF( sFullParams, new_segment_callback ); OFF( sFullParams, new_segment_callback_user_data );
	OFF( sFullParams, encoder_begin_callback ); OFF( sFullParams, encoder_begin_callback_user_data );
	SZ( sSegment ); OFF( sSegment, text ); OFF( sSegment, time ); OFF( sSegment, firstToken ); OFF( sSegment, countTokens );
	SZ( sToken ); OFF( sToken, text ); OFF( sToken, time ); OFF( sToken, probability ); OFF( sToken, probabilityTimestamp ); OFF( sToken, ptsum ); OFF( sToken, vlen );
	OFF( sToken, id ); OFF( sToken, flags );
	SZ( sTimeSpan ); SZ( sTimeInterval ); SZ( sTranscribeLength ); OFF( sTranscribeLength, countSegments ); OFF( sTranscribeLength, countTokens );
	SZ( sModelSetup ); OFF( sModelSetup, impl ); OFF( sModelSetup, flags ); OFF( sModelSetup, adapter );
	SZ( sLoadModelCallbacks ); OFF( sLoadModelCallbacks, progress ); OFF( sLoadModelCallbacks, cancel ); OFF( sLoadModelCallbacks, pv );
	SZ( sLoggerSetup ); OFF( sLoggerSetup, sink ); OFF( sLoggerSetup, context ); OFF( sLoggerSetup, level ); OFF( sLoggerSetup, flags );
	SZ( SpecialTokens ); OFF( SpecialTokens, TranscriptionEnd ); OFF( SpecialTokens, TranscriptionBegin ); OFF( SpecialTokens, TaskTranscribe );
	SZ( sLanguageEntry ); SZ( sLanguageList ); SZ( sProgressSink );
	ENUMV( eFullParamsFlags::Translate ); ENUMV( eFullParamsFlags::NoContext ); ENUMV( eFullParamsFlags::SingleSegment ); ENUMV( eFullParamsFlags::PrintSpecial );
	ENUMV( eFullParamsFlags::PrintProgress ); ENUMV( eFullParamsFlags::PrintRealtime ); ENUMV( eFullParamsFlags::PrintTimestamps );
	ENUMV( eFullParamsFlags::TokenTimestamps ); ENUMV( eFullParamsFlags::SpeedupAudio );
	ENUMV( eResultFlags::Tokens ); ENUMV( eResultFlags::Timestamps ); ENUMV( eResultFlags::NewObject );
	ENUMV( eSamplingStrategy::Greedy ); ENUMV( eSamplingStrategy::BeamSearch );
	ENUMV( eModelImplementation::GPU ); ENUMV( eModelImplementation::Hybrid ); ENUMV( eModelImplementation::Reference );
	ENUMV( eTokenFlags::Special ); ENUMV( eLogLevel::Error ); ENUMV( eLogLevel::Debug ); ENUMV( eSpeakerChannel::NoStereoData );
	guid( "IUnknown", ComLight::IUnknown::iid() ); guid( "iContext", iContext::iid() ); guid( "iModel", iModel::iid() );
	guid( "iTranscribeResult", iTranscribeResult::iid() ); guid( "iAudioBuffer", iAudioBuffer::iid() ); guid( "iAudioReader", iAudioReader::iid() );
	SLOT( iContext, QueryInterface ); SLOT( iContext, AddRef ); SLOT( iContext, Release );
	SLOT( iContext, runFull ); SLOT( iContext, runStreamed ); SLOT( iContext, runCapture ); SLOT( iContext, getResults ); SLOT( iContext, detectSpeaker );
	SLOT( iContext, getModel ); SLOT( iContext, fullDefaultParams ); SLOT( iContext, timingsPrint ); SLOT( iContext, timingsReset );
	SLOT( iModel, createContext ); SLOT( iModel, tokenize ); SLOT( iModel, isMultilingual ); SLOT( iModel, getSpecialTokens ); SLOT( iModel, stringFromToken ); SLOT( iModel, clone );
	SLOT( iTranscribeResult, getSize ); SLOT( iTranscribeResult, getSegments ); SLOT( iTranscribeResult, getTokens );
	SLOT( iAudioBuffer, countSamples ); SLOT( iAudioBuffer, getPcmMono ); SLOT( iAudioBuffer, getPcmStereo ); SLOT( iAudioBuffer, getTime );
	SLOT( iAudioReader, getDuration ); SLOT( iAudioReader, getReader ); SLOT( iAudioReader, requestedStereo );
	// live capture (MfStructs.h, iMediaFoundation.cl.h:28-34)
	SZ( sCaptureParams ); OFF( sCaptureParams, minDuration ); OFF( sCaptureParams, maxDuration ); OFF( sCaptureParams, dropStartSilence ); OFF( sCaptureParams, pauseDuration ); OFF( sCaptureParams, flags );
	SZ( sCaptureCallbacks ); OFF( sCaptureCallbacks, shouldCancel ); OFF( sCaptureCallbacks, captureStatus ); OFF( sCaptureCallbacks, pv );
	ENUMV( eCaptureStatus::Listening ); ENUMV( eCaptureStatus::Voice ); ENUMV( eCaptureStatus::Transcribing ); ENUMV( eCaptureStatus::Stalled ); ENUMV( eCaptureFlags::Stereo );
	{ sCaptureParams d; printf( "  sCaptureParams defaults %g %g %g %g %u\n", d.minDuration, d.maxDuration, d.dropStartSilence, d.pauseDuration, d.flags ); }
	guid( "iAudioCapture", iAudioCapture::iid() );
	SLOT( iAudioCapture, getReader ); SLOT( iAudioCapture, getParams );
	// the media layer (iMediaFoundation.cl.h:36-49, MfStructs.h:5-16)
	SZ( sCaptureDevice ); OFF( sCaptureDevice, displayName ); OFF( sCaptureDevice, endpoint );
	guid( "iMediaFoundation", iMediaFoundation::iid() );
	SLOT( iMediaFoundation, loadAudioFile ); SLOT( iMediaFoundation, openAudioFile ); SLOT( iMediaFoundation, loadAudioFileData );
	SLOT( iMediaFoundation, listCaptureDevices ); SLOT( iMediaFoundation, openCaptureDevice );
	return 0;
}
