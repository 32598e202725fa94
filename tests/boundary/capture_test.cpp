// TEST INFRASTRUCTURE — CPU unit test of the capture loop behind iContext::runCapture (whisper_b200/csrc/capture_loop.h + vad.h, host-only
// code) with a fake transcriber: utterances are cut where the reference's rules cut them (ContextImpl.capture.cpp:193-274), each one is
// the source's samples at the offset it claims, status bits are reported on change, a slow transcriber stalls the loop and costs audio
// instead of memory, the end of the source ends the session with E_EOF.  Run by tests/test_boundary.py; exits 0 on success.
#include "../../whisper_b200/csrc/capture_loop.h"
#include <chrono>
#include <math.h>
#include <stdio.h>
#include <string.h>

static int failures = 0;
#define CHECK( c ) do { if( !( c ) ) { printf( "FAILED line %d: %s\n", __LINE__, #c ); failures++; } } while( 0 )

// quiet noise with "voiced" stretches (amplitude-modulated harmonic tone): [1.0, 2.2) s, [3.4, 8.4) s, [10.0, 10.6) s of a 13 s clip
static std::vector<float> makeSignal()
{
	const size_t n = 16000 * 13;
	std::vector<float> x( n );
	uint32_t lcg = 99;
	for( size_t i = 0; i < n; i++ )
	{
		lcg = lcg * 1664525u + 1013904223u;
		const double t = (double)i / 16000.0;
		double v = 0.001 * ( (double)( lcg >> 8 ) / 8388608.0 - 1.0 );
		const bool voiced = ( t >= 1.0 && t < 2.2 ) || ( t >= 3.4 && t < 8.4 ) || ( t >= 10.0 && t < 10.6 );
		if( voiced )
		{
			double s = 0;
			for( int k = 1; k < 10; k++ ) s += sin( 2 * M_PI * 140.0 * k * t ) / k;
			v += 0.2 * s * 0.5 * ( 1 + sin( 2 * M_PI * 4.0 * t ) );
		}
		x[ i ] = (float)v;
	}
	return x;
}

struct Job { int64_t start; size_t size; bool intact; };

struct Harness
{
	std::vector<float> signal = makeSignal();
	size_t pos = 0;
	int readSleepUs = 0, transcribeSleepMs = 0;
	std::mutex mtx;
	std::vector<Job> jobs;
	std::vector<uint8_t> statuses;
	int transcribeFailAt = -1;

	int32_t read( float* dst, uint32_t cap, uint32_t* got )
	{
		if( readSleepUs ) std::this_thread::sleep_for( std::chrono::microseconds( readSleepUs ) );
		uint32_t n = cap < 777 ? cap : 777;   // a device block that is not a multiple of anything
		if( n > signal.size() - pos ) n = (uint32_t)( signal.size() - pos );
		if( n ) memcpy( dst, signal.data() + pos, (size_t)n * 4 );
		pos += n;
		*got = n;
		return 0;
	}
	int32_t status( uint8_t bits )
	{
		std::lock_guard<std::mutex> lk( mtx );
		statuses.push_back( bits );
		return 0;
	}
	int32_t transcribe( const std::vector<float>& pcm, int64_t first )
	{
		if( transcribeSleepMs ) std::this_thread::sleep_for( std::chrono::milliseconds( transcribeSleepMs ) );
		std::lock_guard<std::mutex> lk( mtx );
		if( (int)jobs.size() == transcribeFailAt ) return (int32_t)0x80004005;
		const bool intact = first >= 0 && (size_t)first + pcm.size() <= signal.size() && 0 == memcmp( pcm.data(), signal.data() + first, pcm.size() * 4 );
		jobs.push_back( Job{ first, pcm.size(), intact } );
		return 0;
	}
	int32_t run( const wsp::CaptureLoop::Settings& st )
	{
		wsp::CaptureLoop loop( [ this ]( float* d, uint32_t c, uint32_t* g ) { return read( d, c, g ); }, [ this ]( uint8_t b ) { return status( b ); },
			[ this ]( const std::vector<float>& p, int64_t f ) { return transcribe( p, f ); }, st );
		int32_t hr = loop.startup();
		while( hr >= 0 ) hr = loop.step();
		CHECK( loop.samplesSeen() == (int64_t)pos );
		return hr;   // the destructor waits for the transcription in flight
	}
};

int main()
{
	using Loop = wsp::CaptureLoop;
	const Loop::Settings defaults = Loop::settingsFromSeconds( 2.0f, 3.0f, 0.25f, 0.333f );   // sCaptureParams' defaults (MfStructs.h:25-32)
	CHECK( defaults.minDuration == 32000 && defaults.maxDuration == 48000 && defaults.dropStartSilence == 4000 && defaults.pauseDuration == 5328 );

	// 1. a transcriber that keeps up (the source is paced so that it always does)
	{
		Harness h;
		h.readSleepUs = 1000;
		CHECK( h.run( defaults ) == Loop::kEndOfStream );
		CHECK( h.jobs.size() >= 3 );
		int64_t prevEnd = 0;
		size_t voicedCovered = 0;
		for( const Job& j : h.jobs )
		{
			CHECK( j.intact );
			CHECK( j.start >= prevEnd );                                  // in stream order, never overlapping
			CHECK( j.size >= defaults.minDuration );                      // nothing shorter than minDuration is handed over ...
			CHECK( j.size < defaults.maxDuration + Loop::kBlock );        // ... and nothing grows past maxDuration (plus the block that crossed it)
			prevEnd = j.start + (int64_t)j.size;
			// how much of the long voiced stretch [3.4 s, 8.4 s) the utterances cover
			const int64_t a = std::max<int64_t>( j.start, 54400 ), b = std::min<int64_t>( prevEnd, 134400 );
			if( b > a ) voicedCovered += (size_t)( b - a );
		}
		CHECK( voicedCovered > 134400 - 54400 - 2 * Loop::kBlock );      // continuous speech is cut into back-to-back pieces, none of it lost
		CHECK( h.jobs.front().start >= 16000 - 8000 && h.jobs.front().start <= 16000 + 2000 );   // leading silence was dropped (dropStartSilence)
		// status reports: Listening first; Voice and Transcribing both seen set and cleared; never Stalled
		CHECK( !h.statuses.empty() && h.statuses.front() == Loop::Listening );
		bool voice = false, transcribing = false, stalled = false;
		for( uint8_t s : h.statuses ) { voice |= ( s & Loop::Voice ) != 0; transcribing |= ( s & Loop::Transcribing ) != 0; stalled |= ( s & Loop::Stalled ) != 0; CHECK( s & Loop::Listening ); }
		CHECK( voice && transcribing && !stalled );
		CHECK( ( h.statuses.back() & ( Loop::Transcribing | Loop::Stalled ) ) == 0 );
		// (reports come from two threads — the transcriber clears its own bit — so their ORDER in the log is not asserted)
	}
	// 2. a transcriber much slower than the source: the loop stalls, audio is dropped, what is handed over is still intact and in order
	{
		Harness h;
		h.transcribeSleepMs = 60;
		CHECK( h.run( defaults ) == Loop::kEndOfStream );
		CHECK( !h.jobs.empty() );
		int64_t prevEnd = 0;
		size_t total = 0;
		for( const Job& j : h.jobs ) { CHECK( j.intact ); CHECK( j.start >= prevEnd ); prevEnd = j.start + (int64_t)j.size; total += j.size; }
		bool stalled = false;
		for( uint8_t s : h.statuses ) stalled |= ( s & Loop::Stalled ) != 0;
		CHECK( stalled );
		CHECK( total < 16000 * 8 );   // well under the ~7 s of voiced audio plus pauses: the rest arrived while stalled
	}
	// 3. a failing transcriber ends the session with its error at the next hand-over
	{
		Harness h;
		h.readSleepUs = 1000;
		h.transcribeFailAt = 1;
		CHECK( h.run( defaults ) == (int32_t)0x80004005 );
		CHECK( h.jobs.size() == 1 );
	}
	// 4. only silence: nothing is ever handed over, the buffer never grows past dropStartSilence + one block
	{
		Harness h;
		for( float& v : h.signal ) v *= 0.0f;
		CHECK( h.run( defaults ) == Loop::kEndOfStream );
		CHECK( h.jobs.empty() );
		CHECK( h.statuses.size() == 1 );
	}
	printf( failures ? "capture_test: %d FAILURES\n" : "capture_test: ok\n", failures );
	return failures ? 1 : 0;
}
