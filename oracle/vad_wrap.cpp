// TEST INFRASTRUCTURE — flat C API over the reference's voice activity detector (Whisper/Whisper/voiceActivityDetection.{h,cpp},
// compiled unmodified from /root/reference by oracle/Makefile into oracle/_ref/liboracle_vad.so).  Only tests/ may load it: it pins
// whisper_b200/csrc/vad.h, the detector behind iContext::runCapture.
#include "stdafx.h"
#include "voiceActivityDetection.h"

extern "C" {
void* ora_vad_create()
{
	Whisper::VAD* v = new Whisper::VAD();
	v->clear();   // the reference's constructor leaves `state` uninitialised; Capture reaches clear() through its first short detect()
	return v;
}
void ora_vad_destroy( void* h ) { delete static_cast<Whisper::VAD*>( h ); }
void ora_vad_clear( void* h ) { static_cast<Whisper::VAD*>( h )->clear(); }
uint64_t ora_vad_detect( void* h, const float* pcm, uint64_t length ) { return static_cast<Whisper::VAD*>( h )->detect( pcm, (size_t)length ); }
}
