// Host side of iContext::runCapture: the loop that listens to a live PCM source, cuts it into utterances with the voice activity
// detector and hands each utterance to the transcriber on a background thread.  Host-only code (no CUDA), unit-tested on the CPU
// with a fake transcriber (tests/boundary/capture_test.cpp).
//
// Reference: Whisper/Whisper/ContextImpl.capture.cpp:75-305 (class Capture).  One step() is one Capture::run():
//   * read one block from the source and run the detector over what has been buffered since the last hand-over;
//   * no voice anywhere in the buffer: once it is longer than dropStartSilence, throw it away (:223-234);
//   * voice, and it reaches to within pauseDuration of the newly read block: keep listening up to maxDuration (:236-244);
//   * voice that ended a while ago: hand over as soon as the buffer holds minDuration (:245-251);
//   * hand-over = post the buffer to the background transcriber if it is idle (:256-262); if it is still busy let the buffer grow
//     to maxDuration, then raise Stalled and DROP incoming audio until the transcriber is free again (:198-216, :264-273).
// The status bits (Listening / Voice / Transcribing / Stalled) are reported through a callback whenever one changes (:98-125), from
// whichever thread changes it — the Transcribing bit is cleared by the transcriber thread, as in the reference.
// The reference posts to the Windows thread pool; here one worker thread is parked on a condition variable.
#pragma once
#include "vad.h"
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace wsp
{
	class CaptureLoop
	{
	public:
		enum Status : uint8_t { Listening = 1, Voice = 2, Transcribing = 4, Stalled = 0x80 };   // eCaptureStatus (MfStructs.h:34-40)
		static constexpr int32_t kEndOfStream = (int32_t)0x80070026;                           // HRESULT_FROM_WIN32( ERROR_HANDLE_EOF ), PcmReader.h:22

		// durations in samples at 16 kHz (CaptureParams, ContextImpl.capture.cpp:56-73)
		struct Settings { uint32_t minDuration, maxDuration, dropStartSilence, pauseDuration; };
		static Settings settingsFromSeconds( float minDuration, float maxDuration, float dropStartSilence, float pauseDuration )
		{
			auto samples = []( float s ) { return (uint32_t)(int32_t)nearbyintf( s * 16000.0f ); };
			return Settings{ samples( minDuration ), samples( maxDuration ), samples( dropStartSilence ), samples( pauseDuration ) };
		}

		using ReadFn = std::function<int32_t( float* dst, uint32_t capacity, uint32_t* written )>;   // *written == 0: the source has ended
		using StatusFn = std::function<int32_t( uint8_t bits )>;                                     // may be empty
		using TranscribeFn = std::function<int32_t( const std::vector<float>& pcm, int64_t firstSample )>;   // called on the worker thread

		static constexpr uint32_t kBlock = 1600;   // samples asked of the source per step (a capture device delivers 10-100 ms at a time)

		CaptureLoop( ReadFn r, StatusFn s, TranscribeFn t, const Settings& st ) : read( std::move( r ) ), report( std::move( s ) ), transcribe( std::move( t ) ), settings( st )
		{
			worker = std::thread( [ this ]() { workerMain(); } );
		}
		~CaptureLoop()
		{
			{
				std::lock_guard<std::mutex> lk( mtx );
				quit = true;
			}
			wake.notify_all();
			worker.join();   // a transcription in flight is allowed to finish (WaitForThreadpoolWorkCallbacks, :157-158)
		}
		CaptureLoop( const CaptureLoop& ) = delete;
		CaptureLoop& operator=( const CaptureLoop& ) = delete;

		int32_t startup() { return setFlag( Listening ); }

		// One iteration.  0 = keep going; negative = failure (the source's, the transcriber's, a callback's, or kEndOfStream).
		int32_t step()
		{
			if( flags.load() & Stalled )
			{
				const int32_t st = workStatus.load();
				if( st < 0 ) return st;
				if( st != 0 ) return readBlock( true );   // still transcribing: the audio that arrives meanwhile is lost
				int32_t hr = clearFlag( Stalled );
				if( hr < 0 ) return hr;
				return post();
			}
			const size_t oldSamples = pcm.size();
			int32_t hr = readBlock( false );
			if( hr < 0 ) return hr;
			const size_t newSamples = pcm.size();
			const size_t lastVoice = vad.detect( pcm.data(), pcm.size() );
			if( lastVoice == 0 )
			{
				clearFlag( Voice );
				if( newSamples < settings.dropStartSilence ) return 0;
				pcm.clear();
				vad.clear();
				pcmStart = nextSample;
				return 0;
			}
			const bool voiceIsRecent = lastVoice + settings.pauseDuration >= oldSamples;
			if( voiceIsRecent )
			{
				setFlag( Voice );
				if( newSamples < settings.maxDuration ) return 0;
			}
			else
			{
				clearFlag( Voice );
				if( newSamples < settings.minDuration ) return 0;
			}
			const int32_t st = workStatus.load();
			if( st < 0 ) return st;
			if( st == 0 ) return post();
			if( newSamples < settings.maxDuration ) return 0;
			setFlag( Stalled );
			return 0;
		}

		uint8_t status() const { return flags.load(); }
		int64_t samplesSeen() const { return nextSample; }

	private:
		ReadFn read;
		StatusFn report;
		TranscribeFn transcribe;
		const Settings settings;
		VoiceDetector vad;
		std::vector<float> pcm;            // audio since the last hand-over (or since the last dropped silence)
		int64_t pcmStart = 0, nextSample = 0;
		std::atomic<uint8_t> flags{ 0 };

		// transcriber thread: workStatus 0 = idle / finished fine, 1 = busy (S_FALSE), negative = failed
		std::thread worker;
		std::mutex mtx;
		std::condition_variable wake;
		std::vector<float> job;
		int64_t jobStart = 0;
		bool jobPending = false, quit = false;
		std::atomic<int32_t> workStatus{ 0 };

		int32_t setFlag( uint8_t bit )
		{
			const uint8_t old = flags.fetch_or( bit );
			if( !report || ( old & bit ) ) return 0;
			return report( (uint8_t)( old | bit ) );
		}
		int32_t clearFlag( uint8_t bit )
		{
			const uint8_t old = flags.fetch_and( (uint8_t)~bit );
			if( !report || !( old & bit ) ) return 0;
			return report( (uint8_t)( old & (uint8_t)~bit ) );
		}

		int32_t readBlock( bool discard )
		{
			float blk[ kBlock ];
			uint32_t got = 0;
			const int32_t hr = read( blk, kBlock, &got );
			if( hr < 0 ) return hr;
			if( got == 0 ) return kEndOfStream;
			if( got > kBlock ) got = kBlock;
			if( !discard ) pcm.insert( pcm.end(), blk, blk + got );
			nextSample += got;
			return 0;
		}

		int32_t post()
		{
			const int32_t hr = setFlag( Transcribing );
			if( hr < 0 ) return hr;
			workStatus.store( 1 );
			{
				std::lock_guard<std::mutex> lk( mtx );
				job.swap( pcm );
				jobStart = pcmStart;
				jobPending = true;
			}
			wake.notify_one();
			pcmStart = nextSample;
			pcm.clear();
			vad.clear();
			return 0;
		}

		void workerMain()
		{
			std::vector<float> mine;
			while( true )
			{
				int64_t start = 0;
				{
					std::unique_lock<std::mutex> lk( mtx );
					wake.wait( lk, [ this ]() { return quit || jobPending; } );
					if( !jobPending ) return;   // quit with nothing queued
					mine.swap( job );
					start = jobStart;
					jobPending = false;
				}
				int32_t st = transcribe( mine, start );
				if( st >= 0 )
				{
					st = clearFlag( Transcribing );
					if( st > 0 ) st = 0;
				}
				workStatus.store( st < 0 ? st : 0 );
			}
		}
	};
}
