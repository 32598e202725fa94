// TEST INFRASTRUCTURE — stands in for Whisper/stdafx.h (the Windows precompiled header) when oracle/Makefile compiles
// Whisper/Whisper/voiceActivityDetection.cpp unmodified with gcc: the C/C++ headers that file relies on, MSVC's std::log10f, and
// DirectX::XMScalarSinCos.  DirectXMath evaluates a minimax polynomial there; sinf / cosf differ from it in the last bit or two of the
// FFT twiddles, which is far below anything the detector's thresholds resolve (stated in tests/test_vad.py).
#pragma once
#include <algorithm>
#include <cmath>
#include <complex>
#include <immintrin.h>
#include <math.h>
#include <memory>
#include <stdint.h>
#include <string.h>
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
namespace std { using ::log10f; }
namespace DirectX
{
	inline void XMScalarSinCos( float* pSin, float* pCos, float value ) { *pSin = sinf( value ); *pCos = cosf( value ); }
}
