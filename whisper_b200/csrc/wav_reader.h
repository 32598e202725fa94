// RIFF/WAVE decoder behind the Linux iMediaFoundation object (initMediaFoundation -> loadAudioFile / openAudioFile /
// loadAudioFileData): 16-bit PCM or 32-bit float, any channel count, any sample rate -> 16 kHz mono f32 (+ optionally the first two
// channels interleaved), delivered sequentially in blocks of any size.  Host-only code (no CUDA), unit-tested on the CPU.
//
// It stands where the reference has Media Foundation (Whisper/MF/loadAudioFile.cpp, AudioBuffer.cpp, PcmReader.cpp): a source reader
// configured for 16 kHz float output, whose samples are down-mixed to mono by averaging the channels (AudioBuffer::appendDownmixedStereo)
// and optionally kept as stereo pairs.  Media Foundation's resampler is not reproducible here; other rates go through linear
// interpolation (out[i] = s[floor(x)] * (1 - t) + s[floor(x) + 1] * t at x = i * rate / 16000), evaluated identically whether the
// file is decoded in one go or block by block.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

namespace wsp
{
	class WavDecoder
	{
	public:
		uint32_t channels = 0, rate = 0, bits = 0;
		uint64_t sourceFrames = 0;        // frames in the data chunk
		std::string error;

		~WavDecoder() { close(); }
		WavDecoder() = default;
		WavDecoder( const WavDecoder& ) = delete;
		WavDecoder& operator=( const WavDecoder& ) = delete;

		bool openFile( const char* path )
		{
			close();
			file = fopen( path, "rb" );
			if( !file ) { error = std::string( "cannot open " ) + path; return false; }
			fseek( file, 0, SEEK_END );
			const long sz = ftell( file );
			fseek( file, 0, SEEK_SET );
			size = sz > 0 ? (uint64_t)sz : 0;
			return parseHeader();
		}
		// the bytes are copied: the reference's loadAudioFileData wraps them in an IMFByteStream that outlives the call the same way
		bool openMemory( const void* data, uint64_t bytes )
		{
			close();
			memory.assign( static_cast<const uint8_t*>( data ), static_cast<const uint8_t*>( data ) + bytes );
			size = bytes;
			return parseHeader();
		}

		// number of 16 kHz samples the stream will deliver in total
		uint64_t outputFrames() const { return rate == 16000 ? sourceFrames : (uint64_t)( (double)sourceFrames * 16000.0 / (double)rate ); }

		// Next `capacity` (or fewer, at the end) output samples: mono[i], and stereo[2i], stereo[2i+1] when `stereo` is not null (for a
		// mono file both carry the mono sample).  Returns the count delivered; 0 = end of stream or a read error (see `error`).
		size_t read( float* mono, float* stereo, size_t capacity )
		{
			size_t done = 0;
			const uint64_t total = outputFrames();
			while( done < capacity && nextOut < total )
			{
				uint64_t i0, i1;
				float t;
				if( rate == 16000 ) { i0 = i1 = nextOut; t = 0.0f; }
				else
				{
					const double x = (double)nextOut * (double)rate / 16000.0;
					i0 = (uint64_t)x;
					i1 = i0 + 1 < sourceFrames ? i0 + 1 : sourceFrames - 1;
					t = (float)( x - (double)i0 );
				}
				if( !ensure( i1 ) ) break;
				const Frame& a = window[ (size_t)( i0 - windowBase ) ];
				const Frame& b = window[ (size_t)( i1 - windowBase ) ];
				if( rate == 16000 )
				{
					mono[ done ] = a.mono;
					if( stereo ) { stereo[ 2 * done ] = a.left; stereo[ 2 * done + 1 ] = a.right; }
				}
				else
				{
					mono[ done ] = a.mono * ( 1.0f - t ) + b.mono * t;
					if( stereo )
					{
						stereo[ 2 * done ] = a.left * ( 1.0f - t ) + b.left * t;
						stereo[ 2 * done + 1 ] = a.right * ( 1.0f - t ) + b.right * t;
					}
				}
				done++;
				nextOut++;
				// frames before i0 of the NEXT output are never needed again
				if( i0 > windowBase + 4096 )
				{
					window.erase( window.begin(), window.begin() + (ptrdiff_t)( i0 - windowBase ) );
					windowBase = i0;
				}
			}
			return done;
		}

	private:
		struct Frame { float mono, left, right; };
		FILE* file = nullptr;
		std::vector<uint8_t> memory;
		uint64_t size = 0, dataOffset = 0, cursor = 0;   // cursor: byte position of the next undecoded source frame
		uint64_t decodedFrames = 0;                      // source frames decoded so far = windowBase + window.size()
		std::vector<Frame> window;                       // decoded source frames [ windowBase, windowBase + window.size() )
		uint64_t windowBase = 0, nextOut = 0;
		std::vector<uint8_t> raw;

		void close()
		{
			if( file ) fclose( file );
			file = nullptr;
			memory.clear();
			window.clear();
			size = dataOffset = cursor = decodedFrames = windowBase = nextOut = 0;
			channels = rate = bits = 0;
			sourceFrames = 0;
		}
		bool bytesAt( uint64_t off, void* dst, size_t n )
		{
			if( off > size || n > size - off ) return false;
			if( file )
			{
				if( fseek( file, (long)off, SEEK_SET ) != 0 ) return false;
				return fread( dst, 1, n, file ) == n;
			}
			memcpy( dst, memory.data() + off, n );
			return true;
		}
		bool parseHeader()
		{
			uint8_t h[ 12 ];
			if( !bytesAt( 0, h, 12 ) || memcmp( h, "RIFF", 4 ) != 0 || memcmp( h + 8, "WAVE", 4 ) != 0 ) { error = "not a RIFF/WAVE stream"; return false; }
			uint32_t format = 0;
			for( uint64_t o = 12; o + 8 <= size; )
			{
				uint8_t ch[ 8 ];
				if( !bytesAt( o, ch, 8 ) ) break;
				const uint32_t len = (uint32_t)ch[ 4 ] | ( (uint32_t)ch[ 5 ] << 8 ) | ( (uint32_t)ch[ 6 ] << 16 ) | ( (uint32_t)ch[ 7 ] << 24 );
				if( memcmp( ch, "fmt ", 4 ) == 0 && len >= 16 )
				{
					uint8_t fm[ 40 ] = {};
					const size_t take = len < sizeof( fm ) ? len : sizeof( fm );
					if( !bytesAt( o + 8, fm, take ) ) break;
					auto u16 = [ & ]( size_t p ) { return (uint32_t)fm[ p ] | ( (uint32_t)fm[ p + 1 ] << 8 ); };
					format = u16( 0 ); channels = u16( 2 ); rate = u16( 4 ) | ( u16( 6 ) << 16 ); bits = u16( 14 );
					if( format == 0xFFFE && len >= 26 ) format = u16( 24 );   // WAVE_FORMAT_EXTENSIBLE: the sub-format's first word
				}
				else if( memcmp( ch, "data", 4 ) == 0 )
				{
					if( !channels || !rate || !( ( format == 1 && bits == 16 ) || ( format == 3 && bits == 32 ) ) )
					{
						error = "only 16-bit PCM and 32-bit float WAV data are supported";
						return false;
					}
					dataOffset = cursor = o + 8;
					const uint64_t avail = size - dataOffset;
					sourceFrames = ( len < avail ? len : avail ) / ( (uint64_t)channels * bits / 8 );
					return true;
				}
				o += 8 + (uint64_t)len + ( len & 1 );
			}
			error = "no audio data found";
			return false;
		}
		// make source frame `index` available in the window
		bool ensure( uint64_t index )
		{
			while( decodedFrames <= index )
			{
				if( decodedFrames >= sourceFrames ) return false;
				const size_t frame = (size_t)channels * bits / 8;
				uint64_t n = sourceFrames - decodedFrames;
				if( n > 8192 ) n = 8192;
				raw.resize( (size_t)n * frame );
				if( !bytesAt( cursor, raw.data(), raw.size() ) ) { error = "read error"; return false; }
				cursor += raw.size();
				for( uint64_t i = 0; i < n; i++ )
				{
					float acc = 0, lr[ 2 ] = { 0, 0 };
					for( uint32_t c = 0; c < channels; c++ )
					{
						const uint8_t* p = raw.data() + (size_t)i * frame + (size_t)c * bits / 8;
						float v;
						if( bits == 16 ) v = (float)(int16_t)( p[ 0 ] | ( p[ 1 ] << 8 ) ) / 32768.0f;
						else memcpy( &v, p, 4 );
						acc += v;
						if( c < 2 ) lr[ c ] = v;
					}
					const float m = acc / (float)channels;
					window.push_back( channels >= 2 ? Frame{ m, lr[ 0 ], lr[ 1 ] } : Frame{ m, m, m } );
				}
				decodedFrames += n;
			}
			return true;
		}
	};
}
