// Front-end + normalisation kernels (HBM/latency-bound, no tensor cores): log-mel spectrogram, window gather, LayerNorm, embedding.
#include "kernels.cuh"
#include "ptx.cuh"
#include <math.h>
#include <stdlib.h>

namespace kern
{
	bool pdlEnabled()
	{
		static const bool on = []() { const char* e = getenv( "WSP_PDL" ); return !( e && e[ 0 ] == '0' ); }();
		return on;
	}

	// float <-> order-preserving int, for atomicMax on floats of any sign
	__device__ __forceinline__ int orderedFromFloat( float f )
	{
		const int i = __float_as_int( f );
		return i >= 0 ? i : i ^ 0x7FFFFFFF;
	}
	__device__ __forceinline__ float floatFromOrdered( int i ) { return __int_as_float( i >= 0 ? i : i ^ 0x7FFFFFFF ); }

	// ---------------------------------------------------------------------------------------------------------------
	// log-mel: one CTA per frame.  Oracle: log_mel_spectrogram (Whisper/source/whisper.cpp:2060-2181):
	//   frame i = hann .* pcm[160*i .. +400) (zero past the end) ; 400-point DFT ; power ; fold bin j with 400-j ;
	//   mel[j] = sum_k folded[k]*filters[j][k] in double ; log10(max(., 1e-10)) -> float.
	// The oracle's recursive f32 FFT (:2009-2056) is replaced by a direct 201-bin DFT with f64 accumulation: exact to f32
	// rounding, so the only difference to the oracle is the oracle's own f32 FFT round-off.
	constexpr int MEL_FFT = 400;
	constexpr int MEL_HOP = 160;
	constexpr int MEL_BINS = 201;
	constexpr int MEL_BANDS = 80;

	__device__ __forceinline__ void melFrame( const MelTables& tb, const float* __restrict__ pcm, int nSamples, int nLen, float* __restrict__ melRaw, int* maxSlot, const int frame )
	{
		__shared__ float sx[ MEL_FFT ];
		__shared__ double2 stw[ MEL_FFT ];          // { cos, sin }( 2 pi m / 400 )
		__shared__ double sA[ 2 ][ MEL_FFT / 4 ];   // [k parity][n]: cosine-side combinations
		__shared__ double sB[ 2 ][ MEL_FFT / 4 ];   //                 sine-side combinations
		__shared__ float sp[ MEL_BINS + 7 ];
		__shared__ float smax[ 8 ];
		const int tid = threadIdx.x;
		const int offset = frame * MEL_HOP;
		for( int n = tid; n < MEL_FFT; n += blockDim.x )
		{
			const int idx = offset + n;
			sx[ n ] = idx < nSamples ? tb.hann[ n ] * pcm[ idx ] : 0.0f;
			stw[ n ] = tb.twiddle[ n ];
		}
		__syncthreads();
		// Real 400-point DFT in a quarter of the multiply-adds.  With w = exp(-2 pi i k / 400):
		//   X[k] = sum_{n<200} y[n] w^n,  y[n] = x[n] + (-1)^k x[n+200]           (w^200 = (-1)^k)
		//        = y[0] + y[100] (-i)^k + sum_{n=1..99} ( y[n] w^n + (-1)^k y[200-n] conj(w^n) )
		//   Re = y[0] + Re(y[100] (-i)^k) + sum cos(t_n) A[n],   A[n] = y[n] + (-1)^k y[200-n]
		//   Im =        Im(y[100] (-i)^k) - sum sin(t_n) B[n],   B[n] = y[n] - (-1)^k y[200-n]
		// (f64 accumulation as before; only the summation order differs from the plain 400-term sum)
		constexpr int Q = MEL_FFT / 4;   // 100
		if( tid < 2 * Q )
		{
			const int par = tid / Q, n = tid - par * Q;
			const double sg = par ? -1.0 : 1.0;
			const double yn = (double)sx[ n ] + sg * (double)sx[ n + 2 * Q ];
			const double ym = n == 0 ? 0.0 : (double)sx[ 2 * Q - n ] + sg * (double)sx[ 4 * Q - n ];   // y[200-n]
			sA[ par ][ n ] = yn + sg * ym;
			sB[ par ][ n ] = yn - sg * ym;
		}
		__syncthreads();
		if( tid < MEL_BINS )
		{
			const int k = tid;
			const int par = k & 1;
			const double sg = par ? -1.0 : 1.0;
			const double y0 = (double)sx[ 0 ] + sg * (double)sx[ 2 * Q ];
			const double y100 = (double)sx[ Q ] + sg * (double)sx[ 3 * Q ];
			double re = y0, im = 0.0;
			switch( k & 3 )   // y[100] * (-i)^k
			{
			case 0: re += y100; break;
			case 1: im -= y100; break;
			case 2: re -= y100; break;
			default: im += y100; break;
			}
			const double* A = sA[ par ];
			const double* Bv = sB[ par ];
			int idx = k;
#pragma unroll 4
			for( int n = 1; n < Q; n++ )
			{
				const double2 w = stw[ idx ];
				re += A[ n ] * w.x;
				im -= Bv[ n ] * w.y;
				idx += k;
				if( idx >= MEL_FFT ) idx -= MEL_FFT;
			}
			const float fr = (float)re, fi = (float)im;
			float pw = fr * fr + fi * fi;
			if( k > 0 && k < MEL_FFT / 2 ) pw = pw + pw;   // fft_out[j] += fft_out[400 - j]  (whisper.cpp:2118-2123)
			sp[ k ] = pw;
		}
		__syncthreads();
		float v = -INFINITY;
		if( tid < MEL_BANDS )
		{
			const float* f = tb.filters + tid * MEL_BINS;
			const short2 range = tb.band[ tid ];   // outside it the filter row is exactly zero: those terms add +0.0
			double sum = 0.0;
			for( int k = range.x; k < range.y; k++ )
				sum += (double)( sp[ k ] * f[ k ] );   // float product, double accumulate (whisper.cpp:2139-2143)
			if( sum < 1e-10 ) sum = 1e-10;
			v = (float)log10( sum );
			melRaw[ (size_t)tid * nLen + frame ] = v;
		}
		// block max -> one atomic
		for( int o = 16; o > 0; o >>= 1 ) v = fmaxf( v, __shfl_xor_sync( 0xffffffffu, v, o ) );
		if( ( tid & 31 ) == 0 ) smax[ tid >> 5 ] = v;
		__syncthreads();
		if( tid == 0 )
		{
			float m = smax[ 0 ];
			for( int w = 1; w < 3; w++ ) m = fmaxf( m, smax[ w ] );   // bands live in warps 0..2
			atomicMax( maxSlot, orderedFromFloat( m ) );
		}
	}

	__global__ void __launch_bounds__( 256 )
		mel_power_kernel( MelTables tb, const float* __restrict__ pcm, int nSamples, int nLen, float* __restrict__ melRaw, int* maxSlot )
	{
		melFrame( tb, pcm, nSamples, nLen, melRaw, maxSlot, blockIdx.x );
	}
	// the same for up to MEL_BATCH chunks in one launch: blockIdx.y = chunk (the driver loop of BASELINE.md §2 transforms B chunks per step)
	__global__ void __launch_bounds__( 256 )
		mel_power_batch_kernel( MelTables tb, MelBatch mb )
	{
		const int b = blockIdx.y;
		if( (int)blockIdx.x >= mb.nLen[ b ] ) return;
		melFrame( tb, mb.pcm[ b ], mb.nSamples[ b ], mb.nLen[ b ], mb.mel[ b ], mb.maxSlots + b, blockIdx.x );
	}

	__device__ __forceinline__ void melNormOne( float* mel, int i, const int* maxSlot )
	{
		// whisper.cpp:2161-2177: mmax (double) -= 8.0; clamp; (x + 4.0) / 4.0 in double, stored as float
		const double mmax = (double)floatFromOrdered( *maxSlot ) - 8.0;
		float v = mel[ i ];
		if( (double)v < mmax ) v = (float)mmax;
		mel[ i ] = (float)( ( (double)v + 4.0 ) / 4.0 );
	}
	__global__ void mel_norm_kernel( float* mel, int count, const int* maxSlot )
	{
		const int i = blockIdx.x * blockDim.x + threadIdx.x;
		if( i >= count ) return;
		melNormOne( mel, i, maxSlot );
	}
	__global__ void mel_norm_batch_kernel( MelBatch mb )
	{
		const int b = blockIdx.y;
		const int i = blockIdx.x * blockDim.x + threadIdx.x;
		if( i >= MEL_BANDS * mb.nLen[ b ] ) return;
		melNormOne( mb.mel[ b ], i, mb.maxSlots + b );
	}

	__global__ void init_max_kernel( int* maxSlot ) { maxSlot[ threadIdx.x ] = orderedFromFloat( -1e20f ); }

	cudaError_t melPower( const MelTables& tb, const float* pcm, int nSamples, int nLen, float* melRaw, int* maxSlot, cudaStream_t s )
	{
		init_max_kernel<<<1, 1, 0, s>>>( maxSlot );
		if( nLen > 0 )
			mel_power_kernel<<<nLen, 256, 0, s>>>( tb, pcm, nSamples, nLen, melRaw, maxSlot );
		return cudaGetLastError();
	}
	cudaError_t melBatch( const MelTables& tb, const MelBatch& mb, cudaStream_t s )
	{
		if( mb.count < 1 || mb.count > MEL_BATCH ) return cudaErrorInvalidValue;
		int maxLen = 0;
		for( int b = 0; b < mb.count; b++ ) maxLen = mb.nLen[ b ] > maxLen ? mb.nLen[ b ] : maxLen;
		init_max_kernel<<<1, mb.count, 0, s>>>( mb.maxSlots );
		if( maxLen > 0 )
		{
			mel_power_batch_kernel<<<dim3( maxLen, mb.count ), 256, 0, s>>>( tb, mb );
			mel_norm_batch_kernel<<<dim3( ( MEL_BANDS * maxLen + 255 ) / 256, mb.count ), 256, 0, s>>>( mb );
		}
		return cudaGetLastError();
	}
	cudaError_t melNormalize( float* mel, int count, const int* maxSlot, cudaStream_t s )
	{
		if( count > 0 )
			mel_norm_kernel<<<( count + 255 ) / 256, 256, 0, s>>>( mel, count, maxSlot );
		return cudaGetLastError();
	}

	// ---------------------------------------------------------------------------------------------------------------
	// window gather: mel[80][nLen] (band-major) -> f16 [1 + frames + 1][80] (time-major) at `offset`, zero beyond nLen
	// (whisper.cpp:1104-1120 copies the slice and zero-pads; the f16 rounding is ggml's conv input conversion, ggml.c:5275-5283)
	__global__ void __launch_bounds__( 256 )
		mel_window_kernel( const float* __restrict__ mel, int nLen, int offset, __half* __restrict__ dst, int frames )
	{
		__shared__ float tile[ MEL_BANDS ][ 33 ];
		const int t0 = blockIdx.x * 32;
		for( int i = threadIdx.x; i < MEL_BANDS * 32; i += blockDim.x )
		{
			const int c = i >> 5, tt = i & 31;
			const int t = t0 + tt;
			const int src = offset + t;
			tile[ c ][ tt ] = ( t < frames && src < nLen ) ? mel[ (size_t)c * nLen + src ] : 0.0f;
		}
		__syncthreads();
		for( int i = threadIdx.x; i < MEL_BANDS * 32; i += blockDim.x )
		{
			const int tt = i / MEL_BANDS, c = i - tt * MEL_BANDS;
			const int t = t0 + tt;
			if( t < frames )
				dst[ (size_t)( t + 1 ) * MEL_BANDS + c ] = __float2half_rn( tile[ c ][ tt ] );
		}
	}
	cudaError_t melWindow( const float* mel, int nLen, int offset, __half* dst, int frames, cudaStream_t s )
	{
		mel_window_kernel<<<( frames + 31 ) / 32, 256, 0, s>>>( mel, nLen, offset, dst, frames );
		return cudaGetLastError();
	}

	// ---------------------------------------------------------------------------------------------------------------
	// LayerNorm (eps 1e-5 inside the sqrt) + affine -> f16: one warp per row, the row stays in registers.
	// Oracle: ggml_compute_forward_norm_f32 (ggml.c:4098-4156) then mul/add with gamma/beta (whisper.cpp:1190-1198);
	// the f16 rounding is the activation conversion the next mul_mat performs (ggml.c:4592-4603).
	constexpr int LN_MAX_V4 = 10;   // d <= 1280

	__device__ __forceinline__ float warpSum( float v )
	{
		for( int o = 16; o > 0; o >>= 1 ) v += __shfl_xor_sync( 0xffffffffu, v, o );
		return v;
	}

	__global__ void __launch_bounds__( 256 )
		layernorm_f16_kernel( const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, __half* __restrict__ out, int rows, int d )
	{
		const int row = blockIdx.x * ( blockDim.x >> 5 ) + ( threadIdx.x >> 5 );
		if( row >= rows ) return;
		const int lane = threadIdx.x & 31;
		const int n4 = d >> 7;   // float4 per lane (d multiple of 128)
		const float4* xr = reinterpret_cast<const float4*>( x + (size_t)row * d );
		float4 v[ LN_MAX_V4 ];
		float sum = 0.0f;
#pragma unroll
		for( int i = 0; i < LN_MAX_V4; i++ )
			if( i < n4 )
			{
				v[ i ] = xr[ i * 32 + lane ];
				sum += v[ i ].x + v[ i ].y + v[ i ].z + v[ i ].w;
			}
		const float mean = warpSum( sum ) / (float)d;
		float sq = 0.0f;
#pragma unroll
		for( int i = 0; i < LN_MAX_V4; i++ )
			if( i < n4 )
			{
				v[ i ].x -= mean; v[ i ].y -= mean; v[ i ].z -= mean; v[ i ].w -= mean;
				sq += v[ i ].x * v[ i ].x + v[ i ].y * v[ i ].y + v[ i ].z * v[ i ].z + v[ i ].w * v[ i ].w;
			}
		const float rstd = 1.0f / sqrtf( warpSum( sq ) / (float)d + 1e-5f );
		const float4* g4 = reinterpret_cast<const float4*>( gamma );
		const float4* b4 = reinterpret_cast<const float4*>( beta );
		uint2* o2 = reinterpret_cast<uint2*>( out + (size_t)row * d );
#pragma unroll
		for( int i = 0; i < LN_MAX_V4; i++ )
			if( i < n4 )
			{
				const float4 g = g4[ i * 32 + lane ];
				const float4 b = b4[ i * 32 + lane ];
				__half2 h0 = __floats2half2_rn( v[ i ].x * rstd * g.x + b.x, v[ i ].y * rstd * g.y + b.y );
				__half2 h1 = __floats2half2_rn( v[ i ].z * rstd * g.z + b.z, v[ i ].w * rstd * g.w + b.w );
				uint2 u;
				u.x = *reinterpret_cast<uint32_t*>( &h0 );
				u.y = *reinterpret_cast<uint32_t*>( &h1 );
				o2[ i * 32 + lane ] = u;
			}
	}
	cudaError_t layerNormF16( const float* x, const float* gamma, const float* beta, __half* out, int rows, int d, cudaStream_t s )
	{
		if( ( d & 127 ) != 0 || d > LN_MAX_V4 * 128 ) return cudaErrorInvalidValue;
		const int warpsPerBlock = 8;
		layernorm_f16_kernel<<<( rows + warpsPerBlock - 1 ) / warpsPerBlock, warpsPerBlock * 32, 0, s>>>( x, gamma, beta, out, rows, d );
		return cudaGetLastError();
	}

	// ---------------------------------------------------------------------------------------------------------------
	// token + positional embedding (whisper.cpp:1536-1548: ggml_get_rows of the f16 embedding, f32 add)
	__global__ void embed_kernel( const __half* __restrict__ te, const float* __restrict__ pe, const int* __restrict__ tokens, const int* __restrict__ dNPast,
		float* __restrict__ x, int N, int d )
	{
		ptx::pdl_launch_dependents();
		ptx::pdl_wait();
		const int col = blockIdx.x;   // b*N + i
		const int i = col % N;
		const int tok = tokens[ col ];
		const int pos = *dNPast + i;
		const __half* src = te + (size_t)tok * d;
		const float* p = pe + (size_t)pos * d;
		float* dst = x + (size_t)col * d;
		for( int e = threadIdx.x; e < d; e += blockDim.x )
			dst[ e ] = __half2float( src[ e ] ) + p[ e ];
	}
	cudaError_t embedTokens( const __half* te, const float* pe, const int* tokens, const int* dNPast, float* x, int B, int N, int d, cudaStream_t s )
	{
		return launchPdl( embed_kernel, dim3( B * N ), dim3( 128 ), 0, s, te, pe, tokens, dNPast, x, N, d );
	}

	__global__ void set_ints_kernel( int* dst, int a, int b ) { dst[ 0 ] = a; dst[ 1 ] = b; }
	cudaError_t setInts( int* dst, int a, int b, cudaStream_t s )
	{
		set_ints_kernel<<<1, 1, 0, s>>>( dst, a, b );
		return cudaGetLastError();
	}
}
