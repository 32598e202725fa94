// TEST INFRASTRUCTURE.  A client of Whisper.dll's public API, compiled ONLY against the reference's own headers
// (-I<reference root>: Whisper/API/*.h + ComLightLib/, nothing from this repository) and linked with libwhisper_b200.so.
// It makes the call sequence of the reference's CLI (Examples/main/main.cpp:210-318):
//   setupLogger -> loadModel -> isMultilingual / getSpecialTokens -> createContext -> fullDefaultParams -> runFull -> getResults
//   -> getSize / getSegments / getTokens -> Release
// with its OWN iAudioBuffer implementation (a COM object defined by the client, consumed by the library), and prints the transcript in
// a line format tests/test_boundary.py compares with the reference's whisper_full fixture.
// With a 7th argument "stream" the clip goes through iContext::runStreamed instead, from the client's OWN iAudioReader object.  The
// reference's headers leave the type that reader hands out incomplete (`struct IMFSourceReader;`, iMediaFoundation.cl.h:5) because
// on Windows it comes from <mfreadwrite.h>; on Linux the client completes it with the pull interface libwhisper_b200.so documents
// (include/whisper_b200_com.h) — the only declaration in this file that does not come from the reference tree.
//   usage: ref_client <model.bin> <pcm.f32> <flags> <language> [calls] [duration_ms] [stream]
#include <string.h>   // the reference headers use strlen without including it (MSVC pulls it in transitively)
#include "Whisper/API/whisperComLight.h"
#include "Whisper/API/sFullParams.h"
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include <vector>
using namespace Whisper;

struct PcmBuffer final : public iAudioBuffer
{
	std::vector<float> pcm;
	uint32_t refs = 1;
	HRESULT COMLIGHTCALL QueryInterface( REFIID riid, void** pp ) override
	{
		if( !pp ) return E_POINTER;
		if( riid == iAudioBuffer::iid() || riid == ComLight::IUnknown::iid() ) { *pp = this; refs++; return S_OK; }
		return E_NOINTERFACE;
	}
	uint32_t COMLIGHTCALL AddRef() override { return ++refs; }
	uint32_t COMLIGHTCALL Release() override { const uint32_t r = --refs; if( !r ) delete this; return r; }
	uint32_t COMLIGHTCALL countSamples() const override { return (uint32_t)pcm.size(); }
	const float* COMLIGHTCALL getPcmMono() const override { return pcm.data(); }
	const float* COMLIGHTCALL getPcmStereo() const override { return nullptr; }
	HRESULT COMLIGHTCALL getTime( int64_t& rdi ) const override { rdi = 0; return S_OK; }
};

// the Linux completion of the reference's forward declaration (see the header comment)
struct IMFSourceReader : public ComLight::IUnknown
{
	virtual HRESULT COMLIGHTCALL readPcm( float* mono, uint32_t capacity, uint32_t* written ) = 0;
};
struct PcmSource final : public IMFSourceReader
{
	const std::vector<float>& pcm;
	size_t pos = 0;
	uint32_t refs = 1;
	explicit PcmSource( const std::vector<float>& p ) : pcm( p ) {}
	HRESULT COMLIGHTCALL QueryInterface( REFIID, void** ) override { return E_NOINTERFACE; }
	uint32_t COMLIGHTCALL AddRef() override { return ++refs; }
	uint32_t COMLIGHTCALL Release() override { const uint32_t r = --refs; if( !r ) delete this; return r; }
	HRESULT COMLIGHTCALL readPcm( float* mono, uint32_t capacity, uint32_t* written ) override
	{
		const size_t n = capacity < 3001 ? capacity : 3001;   // odd-sized blocks
		const size_t take = n < pcm.size() - pos ? n : pcm.size() - pos;
		if( take ) memcpy( mono, pcm.data() + pos, take * 4 );
		pos += take;
		*written = (uint32_t)take;
		return S_OK;
	}
};
struct PcmReaderObj final : public iAudioReader
{
	PcmSource* source;
	uint32_t refs = 1;
	explicit PcmReaderObj( const std::vector<float>& p ) : source( new PcmSource( p ) ) {}
	HRESULT COMLIGHTCALL QueryInterface( REFIID riid, void** pp ) override
	{
		if( !pp ) return E_POINTER;
		if( riid == iAudioReader::iid() || riid == ComLight::IUnknown::iid() ) { *pp = this; refs++; return S_OK; }
		return E_NOINTERFACE;
	}
	uint32_t COMLIGHTCALL AddRef() override { return ++refs; }
	uint32_t COMLIGHTCALL Release() override { const uint32_t r = --refs; if( !r ) { source->Release(); delete this; } return r; }
	HRESULT COMLIGHTCALL getDuration( int64_t& rdi ) const override { rdi = (int64_t)( source->pcm.size() / 160 ) * 100000; return S_OK; }
	HRESULT COMLIGHTCALL getReader( IMFSourceReader** pp ) const override { source->AddRef(); *pp = source; return S_OK; }
	HRESULT COMLIGHTCALL requestedStereo() const override { return S_FALSE; }
};

static int g_segCallbacks = 0;
static HRESULT __cdecl onNewSegment( iContext*, uint32_t n_new, void* ) noexcept { g_segCallbacks += (int)n_new; return S_OK; }
static void __stdcall logSink( void*, eLogLevel lvl, const char* msg ) { if( (int)lvl <= 1 ) fprintf( stderr, "[log %d] %s\n", (int)lvl, msg ); }

#undef CHECK
#define CHECK( expr ) do { const HRESULT _hr = ( expr ); if( FAILED( _hr ) ) { fprintf( stderr, "%s failed: 0x%08x\n", #expr, (unsigned)_hr ); return 2; } } while( 0 )

int main( int argc, char** argv )
{
	if( argc < 5 ) return 1;
	sLoggerSetup ls{};
	ls.sink = &logSink; ls.level = eLogLevel::Warning;
	CHECK( setupLogger( ls ) );
	std::wstring wpath;
	for( const char* p = argv[ 1 ]; *p; p++ ) wpath.push_back( (wchar_t)(unsigned char)*p );
	sModelSetup setup;
	iModel* model = nullptr;
	CHECK( loadModel( wpath.c_str(), setup, nullptr, &model ) );
	SpecialTokens st{};
	CHECK( model->getSpecialTokens( st ) );
	printf( "multilingual %d eot %d sot %d beg %d\n", model->isMultilingual() == S_OK ? 1 : 0, st.TranscriptionEnd, st.TranscriptionStart, st.TranscriptionBegin );
	iContext* ctx = nullptr;
	CHECK( model->createContext( &ctx ) );
	PcmBuffer* buf = new PcmBuffer();
	{
		FILE* f = fopen( argv[ 2 ], "rb" );
		if( !f ) return 3;
		fseek( f, 0, SEEK_END ); const long n = ftell( f ); fseek( f, 0, SEEK_SET );
		buf->pcm.resize( (size_t)n / 4 );
		if( fread( buf->pcm.data(), 4, buf->pcm.size(), f ) != buf->pcm.size() ) return 3;
		fclose( f );
	}
	sFullParams p{};
	CHECK( ctx->fullDefaultParams( eSamplingStrategy::Greedy, &p ) );
	printf( "defaults cpuThreads %d n_max_text_ctx %d max_len %d thold_pt %.3f language 0x%x\n", p.cpuThreads, p.n_max_text_ctx, p.max_len, p.thold_pt, p.language );
	p.flags = (eFullParamsFlags)( (uint32_t)atoi( argv[ 3 ] ) );
	p.language = findLanguageKeyA( argv[ 4 ] );
	p.cpuThreads = 4;
	p.new_segment_callback = &onNewSegment;
	const int calls = argc > 5 ? atoi( argv[ 5 ] ) : 1;
	if( argc > 6 ) p.duration_ms = atoi( argv[ 6 ] );
	if( argc > 7 && 0 == strcmp( argv[ 7 ], "stream" ) )
	{
		PcmReaderObj* reader = new PcmReaderObj( buf->pcm );
		sProgressSink sink{ nullptr, nullptr };
		for( int i = 0; i < calls; i++ )
		{
			reader->source->pos = 0;
			CHECK( ctx->runStreamed( p, sink, reader ) );
		}
		reader->Release();
	}
	else
		for( int i = 0; i < calls; i++ ) CHECK( ctx->runFull( p, buf ) );
	iTranscribeResult* res = nullptr;
	CHECK( ctx->getResults( eResultFlags::Tokens | eResultFlags::Timestamps, &res ) );
	sTranscribeLength len{};
	CHECK( res->getSize( len ) );
	const sSegment* segs = res->getSegments();
	const sToken* toks = res->getTokens();
	printf( "segments %u tokens %u callbacks %d\n", len.countSegments, len.countTokens, g_segCallbacks );
	for( uint32_t i = 0; i < len.countSegments; i++ )
	{
		printf( "seg %lld %lld [", (long long)( segs[ i ].time.begin.ticks / 100000 ), (long long)( segs[ i ].time.end.ticks / 100000 ) );
		for( uint32_t j = 0; j < segs[ i ].countTokens; j++ ) printf( "%s%d", j ? " " : "", toks[ segs[ i ].firstToken + j ].id );
		printf( "] %s\n", segs[ i ].text );
	}
	res->Release();
	buf->Release();
	ctx->Release();
	model->Release();
	return 0;
}
