// whisper_b200_main — command-line transcriber over the COM-style surface of libwhisper_b200.so.
//
// Same role, option names and output formats as the reference's Examples/main (main.cpp:174-353, params.cpp, textWriter.cpp):
//   loadModel -> createContext -> fullDefaultParams -> runFull -> getResults -> txt / srt / vtt next to the input file.
// Inputs are opened through the library's media layer exactly as the reference does (initMediaFoundation -> loadAudioFile, or
// openAudioFile + runStreamed with -st; main.cpp:304-319); on Linux that layer decodes RIFF/WAVE (16-bit PCM or 32-bit float, any rate
// and channel count).  -di labels segments with the louder stereo channel (iContext::detectSpeaker).  Options of the reference that
// have no counterpart here (-owts karaoke script, -su speed-up, colours) are accepted where harmless and rejected otherwise.
#include "whisper_b200_com.h"
#include <algorithm>
#include <math.h>
#include <string.h>
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include <vector>
using namespace Whisper;

namespace
{
	struct Params
	{
		int nThreads = 4, offsetMs = 0, durationMs = 0, maxContext = -1, maxLen = 0, device = 0;
		float wordThold = 0.01f;
		bool translate = false, outTxt = false, outVtt = false, outSrt = false, printSpecial = false, noTimestamps = false, stream = false, diarize = false;
		std::string language = "en", model = "models/ggml-base.en.bin", prompt;
		std::vector<std::string> inputs;
	};

	void usage( const char* exe, const Params& p )
	{
		fprintf( stderr, "\nusage: %s [options] file0.wav file1.wav ...\n\noptions:\n", exe );
		fprintf( stderr, "  -h,       --help          show this help message and exit\n" );
		fprintf( stderr, "  -la,      --list-adapters list the CUDA devices and exit\n" );
		fprintf( stderr, "  -gpu N,   --use-gpu N     [%-7d] CUDA device to use\n", p.device );
		fprintf( stderr, "  -t N,     --threads N     [%-7d] reference thread count whose arithmetic is reproduced (sFullParams::cpuThreads)\n", p.nThreads );
		fprintf( stderr, "  -ot N,    --offset-t N    [%-7d] time offset in milliseconds\n", p.offsetMs );
		fprintf( stderr, "  -d  N,    --duration N    [%-7d] duration of audio to process in milliseconds\n", p.durationMs );
		fprintf( stderr, "  -mc N,    --max-context N [%-7d] maximum number of text context tokens to store\n", p.maxContext );
		fprintf( stderr, "  -ml N,    --max-len N     [%-7d] maximum segment length in characters\n", p.maxLen );
		fprintf( stderr, "  -wt N,    --word-thold N  [%-7.2f] word timestamp probability threshold\n", p.wordThold );
		fprintf( stderr, "  -tr,      --translate     translate from source language to english\n" );
		fprintf( stderr, "  -di,      --diarize       stereo audio diarization: prefix every segment with (speaker 0 / 1 / ?)\n" );
		fprintf( stderr, "  -otxt,    --output-txt    output result in a text file\n" );
		fprintf( stderr, "  -ovtt,    --output-vtt    output result in a vtt file\n" );
		fprintf( stderr, "  -osrt,    --output-srt    output result in a srt file\n" );
		fprintf( stderr, "  -ps,      --print-special print special tokens\n" );
		fprintf( stderr, "  -nt,      --no-timestamps do not print timestamps\n" );
		fprintf( stderr, "  -l LANG,  --language LANG [%-7s] spoken language (\"auto\" = detect)\n", p.language.c_str() );
		fprintf( stderr, "  -m FNAME, --model FNAME   [%-7s] model path\n", p.model.c_str() );
		fprintf( stderr, "  -f FNAME, --file FNAME    path of the input audio file (16-bit or float WAV)\n" );
		fprintf( stderr, "  -st,      --stream        read the file while transcribing it (iContext::runStreamed; not with -ml / -di)\n" );
		fprintf( stderr, "  --prompt TEXT             initial prompt for the model\n\n" );
	}

	void WSPCALL listAdapter( const wchar_t* name, void* ) { printf( "\"%ls\"\n", name ); }

	bool parse( int argc, char** argv, Params& p )
	{
		for( int i = 1; i < argc; i++ )
		{
			const std::string a = argv[ i ];
			auto next = [ & ]() -> const char* { if( i + 1 >= argc ) { fprintf( stderr, "error: %s needs a value\n", a.c_str() ); exit( 1 ); } return argv[ ++i ]; };
			if( a.empty() || a[ 0 ] != '-' ) { p.inputs.push_back( a ); continue; }
			if( a == "-h" || a == "--help" ) { usage( argv[ 0 ], p ); return false; }
			if( a == "-la" || a == "--list-adapters" ) { listGPUs( &listAdapter, nullptr ); return false; }
			else if( a == "-t" || a == "--threads" ) p.nThreads = atoi( next() );
			else if( a == "-ot" || a == "--offset-t" ) p.offsetMs = atoi( next() );
			else if( a == "-d" || a == "--duration" ) p.durationMs = atoi( next() );
			else if( a == "-mc" || a == "--max-context" ) p.maxContext = atoi( next() );
			else if( a == "-ml" || a == "--max-len" ) p.maxLen = atoi( next() );
			else if( a == "-wt" || a == "--word-thold" ) p.wordThold = (float)atof( next() );
			else if( a == "-tr" || a == "--translate" ) p.translate = true;
			else if( a == "-otxt" || a == "--output-txt" ) p.outTxt = true;
			else if( a == "-ovtt" || a == "--output-vtt" ) p.outVtt = true;
			else if( a == "-osrt" || a == "--output-srt" ) p.outSrt = true;
			else if( a == "-ps" || a == "--print-special" ) p.printSpecial = true;
			else if( a == "-nt" || a == "--no-timestamps" ) p.noTimestamps = true;
			else if( a == "-nc" || a == "--no-colors" ) {}
			else if( a == "-l" || a == "--language" ) p.language = next();
			else if( a == "-m" || a == "--model" ) p.model = next();
			else if( a == "-f" || a == "--file" ) p.inputs.push_back( next() );
			else if( a == "-gpu" || a == "--use-gpu" ) p.device = atoi( next() );
			else if( a == "--prompt" ) p.prompt = next();
			else if( a == "-st" || a == "--stream" ) p.stream = true;
			else if( a == "-di" || a == "--diarize" ) p.diarize = true;
			else if( a == "-p" || a == "--processors" || a == "-on" || a == "--offset-n" ) next();
			else { fprintf( stderr, "error: unknown or unsupported argument: %s\n", a.c_str() ); usage( argv[ 0 ], p ); return false; }
		}
		return true;
	}

	// ---- writers: txt / srt / vtt as Examples/main/textWriter.cpp produces them (UTF-8 BOM, CRLF, leading blanks of a segment dropped)
	std::string fmtTime( uint64_t ticks, bool comma )
	{
		const uint64_t ms = ticks / 10000;
		char b[ 64 ];
		snprintf( b, sizeof( b ), "%02d:%02d:%02d%c%03d", (int)( ms / 3600000 ), (int)( ms / 60000 % 60 ), (int)( ms / 1000 % 60 ), comma ? ',' : '.', (int)( ms % 1000 ) );
		return b;
	}
	const char* skipBlank( const char* s ) { while( *s == ' ' || *s == '\t' ) s++; return s; }
	std::string replaceExt( const std::string& path, const char* ext )
	{
		const size_t slash = path.find_last_of( "/\\" ), dot = path.find_last_of( '.' );
		return ( dot != std::string::npos && ( slash == std::string::npos || dot > slash ) ? path.substr( 0, dot ) : path ) + ext;
	}
	enum class Fmt { Txt, Srt, Vtt };
	bool writeResult( iContext* ctx, const std::string& audioPath, Fmt fmt, bool timestamps )
	{
		iTranscribeResult* res = nullptr;
		if( FAILED( ctx->getResults( (eResultFlags)( (uint32_t)eResultFlags::Timestamps | (uint32_t)eResultFlags::Tokens ), &res ) ) ) return false;
		sTranscribeLength len{};
		res->getSize( len );
		const sSegment* segs = res->getSegments();
		const std::string path = replaceExt( audioPath, fmt == Fmt::Txt ? ".txt" : fmt == Fmt::Srt ? ".srt" : ".vtt" );
		FILE* f = fopen( path.c_str(), "wb" );
		if( !f ) { res->Release(); return false; }
		fputs( "\xEF\xBB\xBF", f );
		if( fmt == Fmt::Vtt ) fputs( "WEBVTT\r\n\r\n", f );
		for( uint32_t i = 0; i < len.countSegments; i++ )
		{
			const sSegment& s = segs[ i ];
			const std::string t0 = fmtTime( s.time.begin.ticks, fmt == Fmt::Srt ), t1 = fmtTime( s.time.end.ticks, fmt == Fmt::Srt );
			if( fmt == Fmt::Txt )
			{
				if( timestamps ) fprintf( f, "[%s --> %s]  ", t0.c_str(), t1.c_str() );
				fprintf( f, "%s\r\n", skipBlank( s.text ) );
			}
			else
			{
				if( fmt == Fmt::Srt ) fprintf( f, "%u\r\n", i + 1 );
				fprintf( f, "%s --> %s\r\n%s\r\n\r\n", t0.c_str(), t1.c_str(), skipBlank( s.text ) );
			}
		}
		fclose( f );
		res->Release();
		return true;
	}

	struct SegmentPrinter { bool noTimestamps; bool diarize = false; uint32_t printed = 0; };
	HRESULT onNewSegment( iContext* ctx, uint32_t nNew, void* pv ) noexcept
	{
		SegmentPrinter* sp = static_cast<SegmentPrinter*>( pv );
		iTranscribeResult* res = nullptr;
		if( FAILED( ctx->getResults( eResultFlags::Timestamps, &res ) ) ) return S_OK;
		sTranscribeLength len{};
		res->getSize( len );
		const sSegment* segs = res->getSegments();
		for( uint32_t i = len.countSegments - std::min( nNew, len.countSegments ); i < len.countSegments; i++ )
		{
			// --diarize (Examples/main/main.cpp:95-117): which stereo channel carried this segment
			const char* speaker = "";
			eSpeakerChannel channel;
			if( sp->diarize && SUCCEEDED( ctx->detectSpeaker( segs[ i ].time, channel ) ) && channel != eSpeakerChannel::NoStereoData )
				speaker = channel == eSpeakerChannel::Left ? "(speaker 0)" : ( channel == eSpeakerChannel::Right ? "(speaker 1)" : "(speaker ?)" );
			if( sp->noTimestamps ) printf( "%s", segs[ i ].text );
			else printf( "[%s --> %s]  %s%s\n", fmtTime( segs[ i ].time.begin.ticks, false ).c_str(), fmtTime( segs[ i ].time.end.ticks, false ).c_str(), speaker, segs[ i ].text );
		}
		fflush( stdout );
		return S_OK;
	}
	void WSPCALL collectPrompt( const int* tokens, int n, void* pv ) { static_cast<std::vector<int>*>( pv )->assign( tokens, tokens + n ); }
}

int main( int argc, char** argv )
{
	sLoggerSetup ls;
	ls.flags = eLoggerFlags::UseStandardError;
	ls.level = eLogLevel::Info;
	setupLogger( ls );
	Params params;
	if( !parse( argc, argv, params ) ) return 1;
	if( params.inputs.empty() ) { fprintf( stderr, "error: no input files specified\n" ); usage( argv[ 0 ], params ); return 2; }
	if( params.language != "auto" && findLanguageKeyA( params.language.c_str() ) == UINT32_MAX ) { fprintf( stderr, "error: unknown language '%s'\n", params.language.c_str() ); return 3; }

	std::wstring wmodel( params.model.begin(), params.model.end() );
	const std::wstring adapter = std::to_wstring( params.device );
	sModelSetup setup;
	setup.impl = eModelImplementation::B200;
	setup.adapter = adapter.c_str();
	iModel* model = nullptr;
	HRESULT hr = loadModel( wmodel.c_str(), setup, nullptr, &model );
	if( FAILED( hr ) ) { fprintf( stderr, "failed to load the model: 0x%08x\n", (unsigned)hr ); return 4; }
	std::vector<int> prompt;
	if( !params.prompt.empty() && FAILED( model->tokenize( params.prompt.c_str(), &collectPrompt, &prompt ) ) ) { fprintf( stderr, "failed to tokenize the initial prompt\n" ); return 5; }
	iMediaFoundation* mf = nullptr;
	if( FAILED( initMediaFoundation( &mf ) ) ) { fprintf( stderr, "failed to initialize the media layer\n" ); return 7; }
	iContext* context = nullptr;
	hr = model->createContext( &context );
	if( FAILED( hr ) ) { fprintf( stderr, "failed to initialize whisper context: 0x%08x\n", (unsigned)hr ); return 6; }

	for( const std::string& fname : params.inputs )
	{
		if( model->isMultilingual() == S_FALSE && ( params.language != "en" || params.translate ) )
		{
			params.language = "en";
			params.translate = false;
			fprintf( stderr, "main: WARNING: model is not multilingual, ignoring language and translation options\n" );
		}
		// like the reference's CLI (main.cpp:304-319): the media layer opens the file — streamed (its STREAM_AUDIO build; here opt-in
		// with -st) unless token-level timestamps or diarisation need the whole clip, else loaded into a buffer
		const bool streamed = params.stream && params.maxLen <= 0 && !params.diarize;
		if( params.stream && !streamed ) fprintf( stderr, "main: WARNING: --max-len / --diarize need the whole clip, falling back to buffered mode\n" );
		iAudioBuffer* buffer = nullptr;
		iAudioReader* reader = nullptr;
		hr = streamed ? mf->openAudioFile( fname.c_str(), params.diarize, &reader ) : mf->loadAudioFile( fname.c_str(), params.diarize, &buffer );
		if( FAILED( hr ) ) { fprintf( stderr, "error: cannot read %s (16-bit PCM or 32-bit float RIFF/WAVE expected): 0x%08x\n", fname.c_str(), (unsigned)hr ); return 8; }
		if( params.diarize && buffer && !buffer->getPcmStereo() ) fprintf( stderr, "main: WARNING: %s has one channel, --diarize has nothing to compare\n", fname.c_str() );

		sFullParams wp;
		context->fullDefaultParams( eSamplingStrategy::Greedy, &wp );
		uint32_t flags = (uint32_t)eFullParamsFlags::NoContext;   // several input files are independent clips (main.cpp:262-263)
		if( !params.noTimestamps ) flags |= (uint32_t)eFullParamsFlags::PrintTimestamps;
		if( params.printSpecial ) flags |= (uint32_t)eFullParamsFlags::PrintSpecial;
		if( params.translate ) flags |= (uint32_t)eFullParamsFlags::Translate;
		if( params.maxLen > 0 ) flags |= (uint32_t)eFullParamsFlags::TokenTimestamps;
		wp.flags = (eFullParamsFlags)flags;
		wp.language = params.language == "auto" ? makeLanguageKey( "auto" ) : makeLanguageKey( params.language.c_str() );
		wp.cpuThreads = params.nThreads;
		if( params.maxContext >= 0 ) wp.n_max_text_ctx = params.maxContext;
		wp.offset_ms = params.offsetMs;
		wp.duration_ms = params.durationMs;
		wp.thold_pt = params.wordThold;
		wp.max_len = params.maxLen;
		if( !prompt.empty() ) { wp.prompt_tokens = prompt.data(); wp.prompt_n_tokens = (int)prompt.size(); }
		SegmentPrinter printer{ params.noTimestamps, params.diarize };
		wp.new_segment_callback = &onNewSegment;
		wp.new_segment_callback_user_data = &printer;
		if( streamed )
		{
			hr = context->runStreamed( wp, sProgressSink{ nullptr, nullptr }, reader );
			reader->Release();
		}
		else
		{
			hr = context->runFull( wp, buffer );
			buffer->Release();
		}
		if( FAILED( hr ) ) { fprintf( stderr, "Unable to process audio: 0x%08x\n", (unsigned)hr ); return 10; }
		if( params.noTimestamps ) printf( "\n" );
		if( params.outTxt && !writeResult( context, fname, Fmt::Txt, !params.noTimestamps ) ) fprintf( stderr, "Unable to produce the text file\n" );
		if( params.outSrt && !writeResult( context, fname, Fmt::Srt, true ) ) fprintf( stderr, "Unable to produce the srt file\n" );
		if( params.outVtt && !writeResult( context, fname, Fmt::Vtt, true ) ) fprintf( stderr, "Unable to produce the vtt file\n" );
	}
	context->timingsPrint();
	mf->Release();
	context->Release();
	model->Release();
	return 0;
}
