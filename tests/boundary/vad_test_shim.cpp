// TEST INFRASTRUCTURE — flat C API over whisper_b200/csrc/vad.h (header-only, host-only) so that tests/test_vad.py can drive the product's
// voice detector through ctypes next to the reference's (oracle/_ref/liboracle_vad.so).
#include "../../whisper_b200/csrc/vad.h"

extern "C" {
void* wvad_create() { return new wsp::VoiceDetector(); }
void wvad_destroy( void* h ) { delete static_cast<wsp::VoiceDetector*>( h ); }
void wvad_clear( void* h ) { static_cast<wsp::VoiceDetector*>( h )->clear(); }
uint64_t wvad_detect( void* h, const float* pcm, uint64_t length ) { return static_cast<wsp::VoiceDetector*>( h )->detect( pcm, (size_t)length ); }
void wvad_features( void* h, const float* frame256, float* out3 )
{
	const wsp::VoiceDetector::Features f = static_cast<wsp::VoiceDetector*>( h )->features( frame256 );
	out3[ 0 ] = f.energy; out3[ 1 ] = f.dominantHz; out3[ 2 ] = f.flatness;
}
}
