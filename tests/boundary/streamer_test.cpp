// TEST INFRASTRUCTURE — CPU unit test of the PCM queue behind iContext::runStreamed (whisper_b200/csrc/pcm_streamer.h, host-only code):
// windows handed out at increasing offsets must be exactly the source's samples, whatever the block sizes the source delivers and
// whether or not the background reader is used.  Run by tests/test_boundary.py; exits 0 on success.
#include "../../whisper_b200/csrc/pcm_streamer.h"
#include <stdio.h>
#include <stdlib.h>

static int failures = 0;
#define CHECK( c ) do { if( !( c ) ) { printf( "FAILED line %d: %s\n", __LINE__, #c ); failures++; } } while( 0 )

struct Source
{
	std::vector<float> pcm;
	size_t pos = 0;
	uint32_t lcg = 1, maxBlock = 1;
	bool fixedBlocks = false;
	int failAt = -1, calls = 0;
	int32_t read( float* dst, uint32_t cap, uint32_t* got )
	{
		if( calls++ == failAt ) return (int32_t)0x80004005;
		lcg = lcg * 1664525u + 1013904223u;
		uint32_t want = fixedBlocks ? maxBlock : 1 + ( lcg >> 8 ) % maxBlock;
		if( want > cap ) want = cap;
		if( want > pcm.size() - pos ) want = (uint32_t)( pcm.size() - pos );
		if( want ) memcpy( dst, pcm.data() + pos, (size_t)want * 4 );
		pos += want;
		*got = want;
		return 0;
	}
};

static void scenario( size_t nSamples, uint32_t maxBlock, bool background, const std::vector<size_t>& seeks )
{
	Source src;
	src.pcm.resize( nSamples );
	for( size_t i = 0; i < nSamples; i++ ) src.pcm[ i ] = (float)( ( i * 2654435761u ) & 0xFFFF ) / 65536.0f - 0.5f;
	src.maxBlock = maxBlock;
	const size_t frames = nSamples / 160;
	wsp::PcmStreamer st( [ &src ]( float* d, uint32_t c, uint32_t* g ) { return src.read( d, c, g ); }, frames, background );
	CHECK( st.length() == frames );
	for( size_t seek : seeks )
	{
		const size_t i0 = seek < frames ? seek : frames;
		const size_t i1 = seek + 3000 < frames ? seek + 3000 : frames;
		const float* p = nullptr;
		size_t n = 0;
		const int32_t hr = st.window( i0, i1 - i0, &p, &n );
		CHECK( hr == 0 );
		if( hr != 0 ) return;
		const size_t want = ( i1 - i0 ) * 160 + 240;
		const size_t avail = nSamples > i0 * 160 ? nSamples - i0 * 160 : 0;
		CHECK( n == ( want < avail ? want : avail ) );
		CHECK( st.startFrame() == i0 );
		CHECK( n == 0 || 0 == memcmp( p, src.pcm.data() + i0 * 160, n * 4 ) );
	}
	// same window again is fine, an earlier one is refused like the reference's streamer
	if( !seeks.empty() && seeks.back() > 0 && seeks.back() < frames )
	{
		const float* p; size_t n;
		CHECK( st.window( seeks.back(), 10, &p, &n ) == 0 );
		CHECK( st.window( seeks.back() - 1, 10, &p, &n ) == (int32_t)0x8000FFFF );
	}
}

int main()
{
	for( int bg = 0; bg < 2; bg++ )
	{
		scenario( 16000 * 78, 4096, bg != 0, { 0, 0, 2800, 2801, 5600, 6400, 7700 } );       // the shape of a runStreamed call on 78 s
		scenario( 16000 * 78 + 77, 977, bg != 0, { 1000, 3999, 4000, 7000, 7799, 7800 } );     // ragged tail, start past 0
		scenario( 16000 * 200, 16000, bg != 0, { 0, 3000, 15000, 19990 } );                     // long jump ahead (offset_ms), beyond the prebuffer
		scenario( 16000 * 3 + 5, 7, bg != 0, { 0, 100, 299, 300 } );                           // tiny blocks, window larger than the stream
		scenario( 0, 100, bg != 0, { 0 } );                                                     // empty stream
		scenario( 16000 * 40, 100000, bg != 0, { 0, 5000 } );                                   // seek past the end
	}
	// a failing source surfaces its error, with and without the background reader, and the destructor does not hang
	for( int bg = 0; bg < 2; bg++ )
	{
		Source src;
		src.pcm.assign( 16000 * 100, 0.25f );
		src.maxBlock = 16000;
		src.fixedBlocks = true;   // 31 reads for the first window, the failure comes while the second is being filled
		src.failAt = 40;
		wsp::PcmStreamer st( [ &src ]( float* d, uint32_t c, uint32_t* g ) { return src.read( d, c, g ); }, 10000, bg != 0 );
		const float* p; size_t n;
		CHECK( st.window( 0, 3000, &p, &n ) == 0 );
		CHECK( st.window( 3000, 3000, &p, &n ) == (int32_t)0x80004005 );
	}
	// a streamer that is never asked for anything, with a reader thread parked on a full prebuffer, shuts down cleanly
	{
		Source src;
		src.pcm.assign( 16000 * 200, 0.5f );
		src.maxBlock = 16000;
		wsp::PcmStreamer st( [ &src ]( float* d, uint32_t c, uint32_t* g ) { return src.read( d, c, g ); }, 20000, true );
		const float* p; size_t n;
		CHECK( st.window( 0, 100, &p, &n ) == 0 );
	}
	printf( failures ? "streamer_test: %d FAILURES\n" : "streamer_test: ok\n", failures );
	return failures ? 1 : 0;
}
