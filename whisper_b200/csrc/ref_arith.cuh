// The two pieces of the reference's CPU arithmetic that every decoder kernel has to reproduce operation by operation.
// One definition, shared by the per-op kernels (kernels_decode.cu), the dataflow step kernel (decode_flow.cu) and round 1's barrier
// kernel (decode_mega.cu).
#pragma once
#include <cuda_fp16.h>

namespace kern
{
	// V^T * P with the reference's arithmetic.  In the CPU reference this product takes ggml's "mad" path because V is read
	// transposed (ggml.c:4680-4722): thread ith of nth owns the contiguous key range [ith*dc, (ith+1)*dc), dc = ceil(nkv/nth), and
	// accumulates y[e] = f16( fma( V[j][e], P[j], f32(y[e]) ) ) key by key in an f16 work row (ggml_vec_mad_f16, ggml.c:871-893, with
	// AVX2/F16C: f32 FMA then a round-to-nearest-even store to f16); FINALIZE adds the nth partial rows in f32 in thread order
	// (ggml.c:4613-4640).  With ~1500 keys the f16 running sum swamps the small P[j]*V increments, so the result depends on nth and
	// differs from the exact sum by several percent — that IS the reference's output, and the greedy token sequence depends on it.
	// `refThreads` reproduces it for a given thread count (the reference's default is min(4, cores), whisper.cpp:2605; its own
	// compat shader also "fakes 4 CPU threads", ComputeShaders/mulMatMadMain.hlsl:107).  refThreads = 0 selects exact f32 accumulation.
	__device__ __forceinline__ float pvChainF16( const float* __restrict__ sp, const __half* __restrict__ v, size_t vStride, int j0, int j1 )
	{
		float y = 0.0f;   // always exactly representable in f16
		int j = j0;
		for( ; j + 4 <= j1; j += 4 )
		{
			const float x0 = __half2float( v[ (size_t)j * vStride ] );
			const float x1 = __half2float( v[ (size_t)( j + 1 ) * vStride ] );
			const float x2 = __half2float( v[ (size_t)( j + 2 ) * vStride ] );
			const float x3 = __half2float( v[ (size_t)( j + 3 ) * vStride ] );
			y = __half2float( __float2half_rn( __fmaf_rn( x0, sp[ j ], y ) ) );
			y = __half2float( __float2half_rn( __fmaf_rn( x1, sp[ j + 1 ], y ) ) );
			y = __half2float( __float2half_rn( __fmaf_rn( x2, sp[ j + 2 ], y ) ) );
			y = __half2float( __float2half_rn( __fmaf_rn( x3, sp[ j + 3 ], y ) ) );
		}
		for( ; j < j1; j++ )
			y = __half2float( __float2half_rn( __fmaf_rn( __half2float( v[ (size_t)j * vStride ] ), sp[ j ], y ) ) );
		return y;
	}
	// exp through the reference's f16 table semantics: f32( f16( exp( f16(x) ) ) )  (ggml.c:5075-5077, table :1372-1383)
	__device__ __forceinline__ float expF16Table( float x )
	{
		return __half2float( __float2half_rn( expf( __half2float( __float2half_rn( x ) ) ) ) );
	}
}
