// Voice activity detector behind iContext::runCapture — host-only code (no CUDA), unit-tested on the CPU against the reference's own
// detector compiled from /root/reference (tests/test_vad.py, oracle/_ref/liboracle_vad.so).
//
// Reference: Whisper/Whisper/voiceActivityDetection.{h,cpp} — the algorithm of Moattar & Homayounpour, "A simple but efficient
// real-time voice activity detection algorithm": per 256-sample frame (16 ms) three features — energy, dominant frequency, spectral
// flatness — each compared with a threshold above its running minimum; a frame is speech when two of the three fire.  The detector is
// incremental: detect() is called again and again on a growing buffer and only looks at the frames it has not seen (state.i), and it
// returns the sample position where the most recent speech frame ended (0 = none so far).
//
// Restated, not copied: the reference runs a recursive out-of-place FFT (voiceActivityDetection.cpp:24-49); here it is the same radix-2
// decimation-in-time butterfly network evaluated in place over a bit-reversed input with the twiddles of all stages tabulated once —
// the same butterflies on the same operands, so the spectra agree to float round-off and the decisions agree.
#pragma once
#include <algorithm>
#include <cmath>
#include <complex>
#include <stddef.h>
#include <stdint.h>
#include <vector>

namespace wsp
{
	class VoiceDetector
	{
	public:
		static constexpr uint32_t kFrame = 256;                        // FFT_POINTS
		static constexpr float kBinHz = 16000.0f / (float)kFrame;      // VAD::FFT_STEP

		struct Features { float energy = 0, dominantHz = 0, flatness = 0; };

		VoiceDetector()
		{
			// stage with sub-transforms of size m uses w_m^k = exp( -i * pi * ( k * 2 * n / m ) / n ), the angle rounded exactly as the
			// reference rounds it: ( float(pi) * float(i) ) / float(n) with i = k * 2 * step (voiceActivityDetection.cpp:36-41)
			for( uint32_t m = 2; m <= kFrame; m *= 2 )
				for( uint32_t k = 0; k < m / 2; k++ )
				{
					const uint32_t i = k * 2 * ( kFrame / m );
					const float angle = (float)M_PI * (float)(int)i / (float)(int)kFrame;
					twiddle.emplace_back( cosf( -angle ), sinf( -angle ) );
				}
			for( uint32_t i = 0; i < kFrame; i++ )
			{
				uint32_t r = 0;
				for( uint32_t b = 0; b < 8; b++ ) r |= ( ( i >> b ) & 1u ) << ( 7 - b );
				reversed[ i ] = (uint8_t)r;
			}
			clear();
		}

		// forget everything (VAD::clear): the next detect() starts a new utterance at frame 0
		void clear()
		{
			minimum = Features();
			current = Features();
			lastSpeech = 0;
			silenceRun = 0.0f;
			nextFrame = 0;
		}

		// Look at the frames of pcm[0, length) not seen yet.  Returns the sample index just past the last speech frame, 0 when there is none.
		size_t detect( const float* pcm, size_t length )
		{
			const size_t frames = length / kFrame;
			if( frames == 0 )
			{
				clear();
				return 0;
			}
			for( size_t i = nextFrame; i < frames; i++ )
			{
				current = features( pcm + i * kFrame );
				// the minimum of each feature over the first 30 frames is the noise floor (section 3-3 of the paper)
				if( i == 0 ) minimum = current;
				else if( i < 30 )
				{
					minimum.energy = std::min( minimum.energy, current.energy );
					minimum.dominantHz = std::min( minimum.dominantHz, current.dominantHz );
					minimum.flatness = std::min( minimum.flatness, current.flatness );
				}
				// the energy threshold scales with the logarithm of the floor, the other two are fixed (3-4)
				const float energyThreshold = kEnergyPrim * log10f( minimum.energy );
				int votes = 0;
				if( current.energy - minimum.energy >= energyThreshold ) votes++;
				if( current.dominantHz - minimum.dominantHz >= kDominantPrim ) votes++;
				if( current.flatness - minimum.flatness >= kFlatnessPrim ) votes++;
				if( votes > 1 )
				{
					lastSpeech = ( i + 1 ) * kFrame;
					silenceRun = 0.0f;
				}
				else
				{
					// silence pulls the energy floor towards the current level (3-7)
					silenceRun += 1.0f;
					minimum.energy = ( silenceRun * minimum.energy + current.energy ) / ( silenceRun + 1 );
				}
			}
			nextFrame = std::max( nextFrame, frames );
			return lastSpeech;
		}

		// the three features of one 256-sample frame (samples scaled to the int16 range like the reference, :58, :160)
		Features features( const float* frame ) const
		{
			Features f;
			double sum = 0;
			std::complex<float> x[ kFrame ];
			for( uint32_t j = 0; j < kFrame; j++ )
			{
				float v = frame[ j ];
				v *= 32768.0f;
				x[ reversed[ j ] ] = std::complex<float>( v, 0.0f );
				v *= v;
				sum += v;
			}
			f.energy = sqrtf( (float)( sum * ( 1.0 / kFrame ) ) );

			const std::complex<float>* w = twiddle.data();
			for( uint32_t m = 2; m <= kFrame; m *= 2 )
			{
				const uint32_t half = m / 2;
				for( uint32_t base = 0; base < kFrame; base += m )
					for( uint32_t k = 0; k < half; k++ )
					{
						const std::complex<float> t = w[ k ] * x[ base + k + half ];
						const std::complex<float> a = x[ base + k ];
						x[ base + k ] = a + t;
						x[ base + k + half ] = a - t;
					}
				w += half;
			}

			float best = 0;
			int bestBin = 0;
			for( int i = 0; i < (int)kFrame / 2; i++ )
			{
				const float sq = x[ i ].real() * x[ i ].real() + x[ i ].imag() * x[ i ].imag();
				if( sq <= best ) continue;
				best = sq;
				bestBin = i;
			}
			f.dominantHz = (float)bestBin * kBinHz;

			// spectral flatness: geometric over arithmetic mean of the magnitudes, in dB, sign flipped (:111-125)
			double arithmetic = 0, logs = 0;
			for( uint32_t i = 0; i < kFrame; i++ )
			{
				const float mag = std::abs( x[ i ] );
				arithmetic += mag;
				logs += std::log( mag );
			}
			arithmetic /= kFrame;
			const double geometric = std::exp( logs / kFrame );
			f.flatness = -10.0f * log10f( (float)( geometric / arithmetic ) );
			return f;
		}

		const Features& last() const { return current; }
		const Features& floor() const { return minimum; }

	private:
		static constexpr float kEnergyPrim = 40.0f, kDominantPrim = 185.0f, kFlatnessPrim = 5.0f;   // defaultPrimaryThresholds (:9-16)
		std::vector<std::complex<float>> twiddle;
		uint8_t reversed[ kFrame ];
		Features minimum, current;
		size_t lastSpeech = 0, nextFrame = 0;
		float silenceRun = 0.0f;
	};
}
