// whisper_b200_capture — live transcription from a pipe: the Linux counterpart of the reference's microphone example
// (Examples/MicrophoneCS: iMediaFoundation::openCaptureDevice -> iContext::runCapture).
//
//   arecord -q -f S16_LE -r 16000 -c 1 -t raw | whisper_b200_capture -m ggml-base.en.bin [-l en] [-t 4] [--min 2.0] [--max 3.0]
//
// Raw signed 16-bit little-endian mono PCM at 16 kHz arrives on stdin; createAudioCapture wraps the pipe as the iAudioCapture that
// runCapture listens to.  The library's voice detector cuts utterances (sCaptureParams), transcribes each on a background thread and
// reports segments through new_segment_callback with times relative to the start of the stream; status changes (listening / voice /
// transcribing / stalled) go to stderr.  Ctrl-C or the end of the pipe ends the session.
#include "whisper_b200_com.h"
#include <atomic>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include <vector>
using namespace Whisper;

namespace
{
	std::atomic<bool> g_stop{ false };
	void onSignal( int ) { g_stop = true; }

	HRESULT WSPCALL readStdin( float* mono, uint32_t capacity, uint32_t* written, void* ) noexcept
	{
		// blocks until the pipe has data; 0 samples = the writer closed it
		static std::vector<int16_t> raw;
		raw.resize( capacity );
		const size_t got = fread( raw.data(), sizeof( int16_t ), capacity, stdin );
		for( size_t i = 0; i < got; i++ ) mono[ i ] = (float)raw[ i ] / 32768.0f;
		*written = (uint32_t)got;
		return S_OK;
	}
	HRESULT WSPCALL shouldCancel( void* ) noexcept { return g_stop ? S_FALSE : S_OK; }
	HRESULT WSPCALL onStatus( void*, eCaptureStatus st ) noexcept
	{
		const uint8_t b = (uint8_t)st;
		fprintf( stderr, "[%s%s%s%s ]\n", ( b & 1 ) ? " listening" : "", ( b & 2 ) ? " voice" : "", ( b & 4 ) ? " transcribing" : "", ( b & 0x80 ) ? " STALLED" : "" );
		return S_OK;
	}
	std::string stamp( uint64_t ticks )
	{
		const uint64_t ms = ticks / 10000;
		char buf[ 32 ];
		snprintf( buf, sizeof( buf ), "%02d:%02d:%02d.%03d", (int)( ms / 3600000 ), (int)( ms / 60000 % 60 ), (int)( ms / 1000 % 60 ), (int)( ms % 1000 ) );
		return buf;
	}
	HRESULT onNewSegment( iContext* ctx, uint32_t nNew, void* ) noexcept
	{
		iTranscribeResult* res = nullptr;
		if( FAILED( ctx->getResults( eResultFlags::Timestamps, &res ) ) ) return S_OK;
		sTranscribeLength len{};
		res->getSize( len );
		const sSegment* segs = res->getSegments();
		for( uint32_t i = len.countSegments - ( nNew < len.countSegments ? nNew : len.countSegments ); i < len.countSegments; i++ )
			printf( "[%s --> %s]  %s\n", stamp( segs[ i ].time.begin.ticks ).c_str(), stamp( segs[ i ].time.end.ticks ).c_str(), segs[ i ].text );
		fflush( stdout );
		res->Release();
		return S_OK;
	}
}

int main( int argc, char** argv )
{
	std::string modelPath = "models/ggml-base.en.bin", language = "en";
	int threads = 4, device = 0;
	sCaptureParams cp;
	for( int i = 1; i < argc; i++ )
	{
		const std::string a = argv[ i ];
		auto next = [ & ]() -> const char* { if( i + 1 >= argc ) { fprintf( stderr, "error: %s needs a value\n", a.c_str() ); exit( 1 ); } return argv[ ++i ]; };
		if( a == "-m" || a == "--model" ) modelPath = next();
		else if( a == "-l" || a == "--language" ) language = next();
		else if( a == "-t" || a == "--threads" ) threads = atoi( next() );
		else if( a == "-gpu" || a == "--use-gpu" ) device = atoi( next() );
		else if( a == "--min" ) cp.minDuration = (float)atof( next() );
		else if( a == "--max" ) cp.maxDuration = (float)atof( next() );
		else if( a == "--pause" ) cp.pauseDuration = (float)atof( next() );
		else
		{
			fprintf( stderr, "usage: %s -m model.bin [-l LANG] [-t N] [-gpu N] [--min S] [--max S] [--pause S]  < raw s16le 16 kHz mono PCM\n", argv[ 0 ] );
			return a == "-h" || a == "--help" ? 0 : 1;
		}
	}
	if( findLanguageKeyA( language.c_str() ) == UINT32_MAX ) { fprintf( stderr, "error: unknown language '%s'\n", language.c_str() ); return 3; }
	signal( SIGINT, onSignal );

	sLoggerSetup ls;
	ls.flags = eLoggerFlags::UseStandardError;
	ls.level = eLogLevel::Warning;
	setupLogger( ls );
	const std::wstring wmodel( modelPath.begin(), modelPath.end() ), adapter = std::to_wstring( device );
	sModelSetup setup;
	setup.impl = eModelImplementation::B200;
	setup.adapter = adapter.c_str();
	iModel* model = nullptr;
	HRESULT hr = loadModel( wmodel.c_str(), setup, nullptr, &model );
	if( FAILED( hr ) ) { fprintf( stderr, "failed to load the model: 0x%08x\n", (unsigned)hr ); return 4; }
	iContext* context = nullptr;
	hr = model->createContext( &context );
	if( FAILED( hr ) ) { fprintf( stderr, "failed to initialize whisper context: 0x%08x\n", (unsigned)hr ); return 6; }

	sFullParams wp;
	context->fullDefaultParams( eSamplingStrategy::Greedy, &wp );
	wp.flags = (eFullParamsFlags)( (uint32_t)eFullParamsFlags::NoContext );   // utterances are short: no prompt carry-over between them
	wp.language = makeLanguageKey( language.c_str() );
	wp.cpuThreads = threads;
	wp.new_segment_callback = &onNewSegment;

	iAudioCapture* capture = nullptr;
	if( FAILED( createAudioCapture( &readStdin, nullptr, cp, &capture ) ) ) return 9;
	sCaptureCallbacks callbacks{ &shouldCancel, &onStatus, nullptr };
	hr = context->runCapture( wp, callbacks, capture );
	capture->Release();
	// the pipe running dry ends a session the way an unplugged device does: E_EOF (HRESULT_FROM_WIN32( ERROR_HANDLE_EOF ))
	if( FAILED( hr ) && (uint32_t)hr != 0x80070026u ) { fprintf( stderr, "capture failed: 0x%08x\n", (unsigned)hr ); return 10; }
	context->timingsPrint();
	context->Release();
	model->Release();
	return 0;
}
