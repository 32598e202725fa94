// Decoder per-token kernels (weight/KV streaming, HBM-bound): skinny GEMM, self/cross attention over f16 KV, sampler.
#include "kernels.cuh"
#include "per_device.h"
#include "ptx.cuh"
#include "ref_arith.cuh"
#include <math.h>
#include <cooperative_groups.h>
#include <stdlib.h>

namespace kern
{
	__device__ __forceinline__ float warpSumD( float v )
	{
		for( int o = 16; o > 0; o >>= 1 ) v += __shfl_xor_sync( 0xffffffffu, v, o );
		return v;
	}
	__device__ __forceinline__ float warpMaxD( float v )
	{
		for( int o = 16; o > 0; o >>= 1 ) v = fmaxf( v, __shfl_xor_sync( 0xffffffffu, v, o ) );
		return v;
	}

	// ---------------------------------------------------------------------------------------------------------------
	// Skinny GEMM: out[col][n] = sum_k W[n][k] * x[col][k] for a handful of columns (decoder tokens x chunks).
	//
	// Replaces mulMatByRowTiled.hlsl (32 % of the reference's GPU time, ComputeShaders/mulMatByRowTiled.hlsl:42-169) and the N=1
	// case of ggml_compute_forward_mul_mat_f16_f32 (ggml.c:4447-4749).  The op is pure weight streaming, so the design goal is
	// HBM rate: each lane issues 16-byte loads covering 64 contiguous bytes of a weight row per quad, several in flight; the f16
	// activations (optionally LayerNorm-ed on the fly) sit in shared memory; products go through mma.sync.m16n8k16 (f16 x f16
	// -> f32, the oracle's arithmetic) because 8..16 columns exactly fill the n dimension and the tensor pipe is idle otherwise.
	// The k index is permuted identically for A and B fragments so that every fragment load is one 128-bit access.
	//
	// CTA = 16 weight rows x (up to 16 columns); 8 warps split K; cross-warp reduction through shared memory; fused epilogue.
	constexpr int SK_ROWS = 16;
	constexpr int SK_COLS = 16;      // columns per CTA pass (2 n-tiles of 8)
	constexpr int SK_WARPS = 8;
	constexpr int SK_PAD = 32;       // halves of padding per activation row: row stride = 64 (mod 128) bytes -> conflict-free LDS.128

	__device__ __forceinline__ void mma16816( float* c, const uint32_t* a, uint32_t b0, uint32_t b1 )
	{
		asm volatile(
			"mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
			: "+f"( c[ 0 ] ), "+f"( c[ 1 ] ), "+f"( c[ 2 ] ), "+f"( c[ 3 ] )
			: "r"( a[ 0 ] ), "r"( a[ 1 ] ), "r"( a[ 2 ] ), "r"( a[ 3 ] ), "r"( b0 ), "r"( b1 ) );
	}

	__device__ __forceinline__ uint4 ldg_stream( const uint4* p )
	{
		uint4 r;
		asm volatile( "ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"( r.x ), "=r"( r.y ), "=r"( r.z ), "=r"( r.w ) : "l"( p ) );
		return r;
	}

	__global__ void __launch_bounds__( SK_WARPS * 32 )
		skinny_gemm_kernel( SkinnyArgs a )
	{
		extern __shared__ __align__( 16 ) uint8_t sk_smem[];
		const int K = a.K;
		const int rowStride = K + SK_PAD;   // halves
		__half* sx = reinterpret_cast<__half*>( sk_smem );                       // [SK_COLS][rowStride]
		float* red = reinterpret_cast<float*>( sx + (size_t)SK_COLS * rowStride ); // [SK_WARPS][SK_ROWS][SK_COLS]

		const int tid = threadIdx.x;
		const int warp = tid >> 5;
		const int lane = tid & 31;
		const int row0 = blockIdx.x * SK_ROWS;
		const int col0 = blockIdx.y * SK_COLS;
		const int ncols = min( SK_COLS, a.nCols - col0 );
		ptx::pdl_launch_dependents();

		// ---- weight stream: the first batch of 16-byte loads is issued before anything else, so HBM latency overlaps the
		//      activation staging below (weights do not depend on the previous kernel's output) ----
		const int g = lane >> 2, t = lane & 3;
		const int rA = min( row0 + g, a.nOut - 1 );
		const int rB = min( row0 + g + 8, a.nOut - 1 );
		const uint4* wA = reinterpret_cast<const uint4*>( a.W + (size_t)rA * K + 8 * t );
		const uint4* wB = reinterpret_cast<const uint4*>( a.W + (size_t)rB * K + 8 * t );
		const int steps = K / 32;
		constexpr int U = 8;
		uint4 va[ U ], vb[ U ];
#pragma unroll
		for( int u = 0; u < U; u++ )
		{
			const int st = warp + u * SK_WARPS;
			if( st < steps )
			{
				va[ u ] = ldg_stream( wA + st * 4 );   // 32 halves = 4 uint4 per step
				vb[ u ] = ldg_stream( wB + st * 4 );
			}
		}

		// everything below reads what the previous kernel wrote
		ptx::pdl_wait();

		// ---- stage activations as f16 (LayerNorm fused when gamma is given) ----
		for( int c = warp; c < SK_COLS; c += SK_WARPS )
		{
			__half* dst = sx + (size_t)c * rowStride;
			if( c >= ncols )
			{
				// unused columns of a tile that is computed must be finite; the second n-tile is skipped entirely when ncols <= 8
				if( c < 8 || ncols > 8 )
				{
					uint4* z4 = reinterpret_cast<uint4*>( dst );
					for( int k = lane; k < K / 8; k += 32 ) z4[ k ] = make_uint4( 0, 0, 0, 0 );
				}
				continue;
			}
			if( a.xF32 )
			{
				const float* src = a.xF32 + (size_t)( col0 + c ) * a.xStride;
				if( a.gamma )
				{
					// oracle: ggml_norm (ggml.c:4098-4156) + gamma/beta, then f16 conversion at the mul_mat (ggml.c:4592-4603).
					// The row (K = d <= 1280, multiple of 128) is loaded once into registers, all loads in flight together.
					constexpr int MAXV = 10;
					const int n4 = K >> 7;
					const float4* s4 = reinterpret_cast<const float4*>( src );
					float4 v[ MAXV ];
					float s = 0.0f;
#pragma unroll
					for( int q4 = 0; q4 < MAXV; q4++ )
						if( q4 < n4 )
						{
							v[ q4 ] = s4[ q4 * 32 + lane ];
							s += v[ q4 ].x + v[ q4 ].y + v[ q4 ].z + v[ q4 ].w;
						}
					const float mean = warpSumD( s ) / (float)K;
					float sq = 0.0f;
#pragma unroll
					for( int q4 = 0; q4 < MAXV; q4++ )
						if( q4 < n4 )
						{
							v[ q4 ].x -= mean; v[ q4 ].y -= mean; v[ q4 ].z -= mean; v[ q4 ].w -= mean;
							sq += v[ q4 ].x * v[ q4 ].x + v[ q4 ].y * v[ q4 ].y + v[ q4 ].z * v[ q4 ].z + v[ q4 ].w * v[ q4 ].w;
						}
					const float rstd = 1.0f / sqrtf( warpSumD( sq ) / (float)K + 1e-5f );
					const float4* g4 = reinterpret_cast<const float4*>( a.gamma );
					const float4* b4 = reinterpret_cast<const float4*>( a.beta );
					uint2* d2 = reinterpret_cast<uint2*>( dst );
#pragma unroll
					for( int q4 = 0; q4 < MAXV; q4++ )
						if( q4 < n4 )
						{
							const float4 g = g4[ q4 * 32 + lane ];
							const float4 bb = b4[ q4 * 32 + lane ];
							__half2 h0 = __floats2half2_rn( v[ q4 ].x * rstd * g.x + bb.x, v[ q4 ].y * rstd * g.y + bb.y );
							__half2 h1 = __floats2half2_rn( v[ q4 ].z * rstd * g.z + bb.z, v[ q4 ].w * rstd * g.w + bb.w );
							uint2 u;
							u.x = *reinterpret_cast<uint32_t*>( &h0 );
							u.y = *reinterpret_cast<uint32_t*>( &h1 );
							d2[ q4 * 32 + lane ] = u;
						}
				}
				else
					for( int k = lane; k < K; k += 32 ) dst[ k ] = __float2half_rn( src[ k ] );
			}
			else
			{
				const uint4* src = reinterpret_cast<const uint4*>( a.xF16 + (size_t)( col0 + c ) * a.xStride );
				uint4* d4 = reinterpret_cast<uint4*>( dst );
				for( int k = lane; k < K / 8; k += 32 ) d4[ k ] = src[ k ];
			}
		}
		__syncthreads();

		// ---- main loop: 32 k per step, steps interleaved across the 8 warps (8 warps x 64 B = 512 contiguous bytes per row) ----
		const __half* xb0 = sx + (size_t)g * rowStride + 8 * t;         // n-tile 0: column g
		const __half* xb1 = sx + (size_t)( g + 8 ) * rowStride + 8 * t; // n-tile 1: column g+8
		float acc0[ 4 ] = { 0, 0, 0, 0 }, acc1[ 4 ] = { 0, 0, 0, 0 };
		const bool twoTiles = ncols > 8;
		for( int s0 = warp; s0 < steps; )
		{
#pragma unroll
			for( int u = 0; u < U; u++ )
			{
				const int st = s0 + u * SK_WARPS;
				if( st < steps )
				{
					const uint4 x0 = *reinterpret_cast<const uint4*>( xb0 + st * 32 );
					const uint32_t a1[ 4 ] = { va[ u ].x, vb[ u ].x, va[ u ].y, vb[ u ].y };
					const uint32_t a2[ 4 ] = { va[ u ].z, vb[ u ].z, va[ u ].w, vb[ u ].w };
					mma16816( acc0, a1, x0.x, x0.y );
					mma16816( acc0, a2, x0.z, x0.w );
					if( twoTiles )
					{
						const uint4 x1 = *reinterpret_cast<const uint4*>( xb1 + st * 32 );
						mma16816( acc1, a1, x1.x, x1.y );
						mma16816( acc1, a2, x1.z, x1.w );
					}
				}
			}
			s0 += SK_WARPS * U;
			if( s0 < steps )
			{
#pragma unroll
				for( int u = 0; u < U; u++ )
				{
					const int st = s0 + u * SK_WARPS;
					if( st < steps )
					{
						va[ u ] = ldg_stream( wA + st * 4 );
						vb[ u ] = ldg_stream( wB + st * 4 );
					}
				}
			}
		}

		// ---- cross-warp reduction ----
		// c0,c1: (row g, cols 2t, 2t+1); c2,c3: (row g+8, same cols)
		float* my = red + (size_t)warp * SK_ROWS * SK_COLS;
		my[ g * SK_COLS + 2 * t + 0 ] = acc0[ 0 ];
		my[ g * SK_COLS + 2 * t + 1 ] = acc0[ 1 ];
		my[ ( g + 8 ) * SK_COLS + 2 * t + 0 ] = acc0[ 2 ];
		my[ ( g + 8 ) * SK_COLS + 2 * t + 1 ] = acc0[ 3 ];
		my[ g * SK_COLS + 8 + 2 * t + 0 ] = acc1[ 0 ];
		my[ g * SK_COLS + 8 + 2 * t + 1 ] = acc1[ 1 ];
		my[ ( g + 8 ) * SK_COLS + 8 + 2 * t + 0 ] = acc1[ 2 ];
		my[ ( g + 8 ) * SK_COLS + 8 + 2 * t + 1 ] = acc1[ 3 ];
		__syncthreads();

		// ---- epilogue: thread -> (col = tid / 16, row = tid % 16): consecutive threads write consecutive output features ----
		const int c = tid >> 4, r = tid & 15;
		const int n = row0 + r;
		if( c >= ncols || n >= a.nOut ) return;
		float v = 0.0f;
#pragma unroll
		for( int w = 0; w < SK_WARPS; w++ ) v += red[ ( (size_t)w * SK_ROWS + r ) * SK_COLS + c ];
		const int col = col0 + c;
		switch( a.epi )
		{
		case SK_QKV:
		{
			const int which = n / a.d;
			const int nn = n - which * a.d;
			if( which == 0 )
				a.outF32[ (size_t)col * a.ld + nn ] = ( v + a.bias[ n ] ) * a.scale;
			else
			{
				const int b = col / a.N, i = col - b * a.N;
				// self-KV cache rows are head-major [b][h][pos][64]: one head's rows are contiguous (bulk-copied by decode_flow.cu)
				const int pos = min( *a.dNPast + i, a.nTextCtx - 1 );
				const size_t off = ( ( (size_t)b * ( a.d >> 6 ) + ( nn >> 6 ) ) * a.nTextCtx + pos ) * 64 + ( nn & 63 );
				if( which == 1 ) a.kCache[ off ] = __float2half_rn( v * a.scale );
				else a.vCache[ off ] = __float2half_rn( v + a.bias[ n ] );
			}
			break;
		}
		case SK_BIAS_RESID:
		{
			float* p = a.outF32 + (size_t)col * a.ld + n;
			*p = v + a.bias[ n ] + *p;
			break;
		}
		case SK_Q_SCALE:
			a.outF32[ (size_t)col * a.ld + n ] = ( v + a.bias[ n ] ) * a.scale;
			break;
		case SK_GELU_F16:
			a.outF16[ (size_t)col * a.ld + n ] = __float2half_rn( ptx::gelu_f16_semantics( v + a.bias[ n ] ) );
			break;
		default:
			a.outF32[ (size_t)col * a.ld + n ] = v;
			break;
		}
	}

	static size_t skinnySmem( int K ) { return (size_t)SK_COLS * ( K + SK_PAD ) * sizeof( __half ) + (size_t)SK_WARPS * SK_ROWS * SK_COLS * sizeof( float ); }
	static PerDeviceMax g_skinnyAttr;
	static cudaError_t crossPrepare( int T );
	cudaError_t prepare( int maxK )
	{
		cudaError_t ce = crossPrepare( 1500 );
		if( ce != cudaSuccess ) return ce;
		return g_skinnyAttr.raise( skinnySmem( maxK ), []( size_t n ) { return cudaFuncSetAttribute( skinny_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)n ); } );
	}
	cudaError_t skinnyGemm( const SkinnyArgs& a, cudaStream_t s )
	{
		if( a.K % 32 != 0 ) return cudaErrorInvalidValue;
		if( a.gamma && ( a.K % 128 != 0 || a.K > 1280 ) ) return cudaErrorInvalidValue;   // fused LayerNorm keeps the row in registers
		const size_t smem = skinnySmem( a.K );
		{
			cudaError_t e = prepare( a.K );
			if( e != cudaSuccess ) return e;
		}
		dim3 grid( ( a.nOut + SK_ROWS - 1 ) / SK_ROWS, ( a.nCols + SK_COLS - 1 ) / SK_COLS );
		return launchPdl( skinny_gemm_kernel, grid, dim3( SK_WARPS * 32 ), smem, s, a );
	}

	// ---------------------------------------------------------------------------------------------------------------
	// V^T * P and the softmax exponential with the reference's arithmetic: pvChainF16 / expF16Table in ref_arith.cuh
	constexpr int PV_MAX_THREADS = 16;

	// ---------------------------------------------------------------------------------------------------------------
	// Decoder self-attention, one CTA per (chunk, head, query).  Oracle: whisper.cpp:1616-1661 — K*Q (Q rounded to f16 by the
	// mul_mat, ggml.c:4599), causal mask, softmax, V^T * P.  KV cache rows are head-major [b][h][pos][64] f16.
	constexpr int SA_THREADS = 128;
	constexpr int SA_MAXKV = 448;

	__global__ void __launch_bounds__( SA_THREADS )
		self_attn_decode_kernel( const float* __restrict__ q, const __half* __restrict__ kCache, const __half* __restrict__ vCache, __half* __restrict__ out,
			int N, int H, int d, int nTextCtx, const int* __restrict__ dNPast, int refThreads )
	{
		__shared__ float sq[ 64 ];
		__shared__ float sp[ SA_MAXKV ];
		__shared__ float sred[ SA_THREADS / 32 ];
		__shared__ float so[ PV_MAX_THREADS ][ 64 ];
		ptx::pdl_launch_dependents();
		ptx::pdl_wait();
		const int bh = blockIdx.x;
		const int b = bh / H, h = bh - b * H;
		const int i = blockIdx.y;
		const int tid = threadIdx.x;
		const int nkv = min( *dNPast + i + 1, nTextCtx );   // causal: keys visible to query i
		const int nkvAll = min( *dNPast + N, nTextCtx );     // columns of the KQ matrix the reference partitions over its threads
		const int col = b * N + i;
		if( tid < 64 )
			sq[ tid ] = __half2float( __float2half_rn( q[ (size_t)col * d + h * 64 + tid ] ) );
		__syncthreads();
		const __half* kb = kCache + ( (size_t)b * H + h ) * nTextCtx * 64;
		const __half* vb = vCache + ( (size_t)b * H + h ) * nTextCtx * 64;
		float lmax = -INFINITY;
		for( int j = tid; j < nkv; j += SA_THREADS )
		{
			const uint4* kr = reinterpret_cast<const uint4*>( kb + (size_t)j * 64 );
			float s = 0.0f;
#pragma unroll
			for( int c = 0; c < 8; c++ )
			{
				const uint4 u = kr[ c ];
				const __half2* h2 = reinterpret_cast<const __half2*>( &u );
#pragma unroll
				for( int e = 0; e < 4; e++ )
				{
					const float2 f = __half22float2( h2[ e ] );
					s += f.x * sq[ c * 8 + e * 2 ] + f.y * sq[ c * 8 + e * 2 + 1 ];
				}
			}
			sp[ j ] = s;
			lmax = fmaxf( lmax, s );
		}
		lmax = warpMaxD( lmax );
		if( ( tid & 31 ) == 0 ) sred[ tid >> 5 ] = lmax;
		__syncthreads();
		float mx = sred[ 0 ];
		for( int w = 1; w < SA_THREADS / 32; w++ ) mx = fmaxf( mx, sred[ w ] );
		__syncthreads();
		float lsum = 0.0f;
		for( int j = tid; j < nkv; j += SA_THREADS )
		{
			const float e = expF16Table( sp[ j ] - mx );
			sp[ j ] = e;
			lsum += e;
		}
		lsum = warpSumD( lsum );
		if( ( tid & 31 ) == 0 ) sred[ tid >> 5 ] = lsum;
		__syncthreads();
		float tot = 0.0f;
		for( int w = 0; w < SA_THREADS / 32; w++ ) tot += sred[ w ];
		const float inv = 1.0f / tot;
		for( int j = tid; j < nkv; j += SA_THREADS ) sp[ j ] *= inv;   // the reference normalises P before V^T*P (ggml.c:5085-5090)
		__syncthreads();
		const int e = tid & 63;
		if( refThreads > 0 )
		{
			// reference arithmetic: refThreads f16 chains over contiguous key ranges, summed in f32 in thread order
			const int dc = ( nkvAll + refThreads - 1 ) / refThreads;
			for( int part = tid >> 6; part < refThreads; part += SA_THREADS / 64 )
			{
				const int j0 = min( part * dc, nkv ), j1 = min( ( part + 1 ) * dc, nkv );   // masked keys contribute exactly 0
				so[ part ][ e ] = pvChainF16( sp, vb + e, (size_t)64, j0, j1 );
			}
			__syncthreads();
			if( tid < 64 )
			{
				float acc = so[ 0 ][ tid ];
				for( int k = 1; k < refThreads; k++ ) acc += so[ k ][ tid ];
				out[ (size_t)col * d + h * 64 + tid ] = __float2half_rn( acc );
			}
		}
		else
		{
			const int half = tid >> 6;
			float o = 0.0f;
			for( int j = half; j < nkv; j += 2 )
				o += sp[ j ] * __half2float( vb[ (size_t)j * 64 + e ] );
			so[ half ][ e ] = o;
			__syncthreads();
			if( tid < 64 )
				out[ (size_t)col * d + h * 64 + tid ] = __float2half_rn( so[ 0 ][ tid ] + so[ 1 ][ tid ] );
		}
	}
	cudaError_t selfAttnDecode( const float* q, const __half* kCache, const __half* vCache, __half* out, int B, int N, int H, int d, int nTextCtx, const int* dNPast, int refThreads, cudaStream_t s )
	{
		if( nTextCtx > SA_MAXKV || refThreads < 0 || refThreads > PV_MAX_THREADS ) return cudaErrorInvalidValue;
		dim3 grid( B * H, N );
		return launchPdl( self_attn_decode_kernel, grid, dim3( SA_THREADS ), 0, s, q, kCache, vCache, out, N, H, d, nTextCtx, dNPast, refThreads );
	}

	// ---------------------------------------------------------------------------------------------------------------
	// Decoder cross-attention, one CTA per (chunk, head, query); K and V memories are [b][h][T][64] f16, written by the
	// encoder's cross-KV GEMM epilogue.  Oracle: whisper.cpp:1688-1749 (no mask, K pre-scaled at encode time :1465).
	// The per-sequence 2*T*d bytes per layer are the dominant per-chunk HBM traffic of a decode step:
	// every access below is a 16-byte load, 8 lanes per 128-byte row, fully coalesced.
	constexpr int CA_THREADS = 256;
	constexpr int CA_MAXT = 1536;
	constexpr int CA_UNROLL = 12;

	__device__ __forceinline__ void cp_async16( void* smemDst, const void* gmemSrc )
	{
		asm volatile( "cp.async.cg.shared.global [%0], [%1], 16;" ::"r"( ptx::smem_u32( smemDst ) ), "l"( gmemSrc ) : "memory" );
	}
	__device__ __forceinline__ void cp_async_commit() { asm volatile( "cp.async.commit_group;" ::: "memory" ); }
	__device__ __forceinline__ void cp_async_wait_all() { asm volatile( "cp.async.wait_group 0;" ::: "memory" ); }

	__global__ void __launch_bounds__( CA_THREADS )
		cross_attn_decode_kernel( const float* __restrict__ q, const __half* __restrict__ kMem, const __half* __restrict__ vMem, __half* __restrict__ out,
			int N, int H, int d, int T, int refThreads )
	{
		extern __shared__ __align__( 16 ) uint8_t ca_smem[];   // the whole V tile of this (chunk, head): T x 128 bytes
		__shared__ float sp[ CA_MAXT ];
		__shared__ float sred[ CA_THREADS / 32 ];
		__shared__ float so[ CA_THREADS / 32 ][ 4 ][ 64 ];   // per warp, per row-group partial outputs (8 KB) | [refThreads][64]
		const int bh = blockIdx.x;
		const int b = bh / H, h = bh - b * H;
		const int i = blockIdx.y;
		const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
		const int col = b * N + i;
		const int sub = lane & 7;        // which 16-byte chunk of the 128-byte row
		const int rgrp = lane >> 3;      // row within a group of 4
		const uint4* K4 = reinterpret_cast<const uint4*>( kMem + (size_t)bh * T * 64 );
		const uint4* V4 = reinterpret_cast<const uint4*>( vMem + (size_t)bh * T * 64 );

		ptx::pdl_launch_dependents();
		// V streams into shared memory (cp.async, no registers) behind the whole K phase.  Neither V nor K depends on the previous
		// kernel (the cross-KV memories were written by the encoder), so both prefetches are issued before the dependency wait.
		{
			uint4* sv = reinterpret_cast<uint4*>( ca_smem );
			for( int idx = tid; idx < T * 8; idx += CA_THREADS ) cp_async16( sv + idx, V4 + idx );
			cp_async_commit();
		}
		constexpr int JSTEP = ( CA_THREADS / 32 ) * 4;
		int jb = warp * 4 + rgrp;
		uint4 u[ CA_UNROLL ];
#pragma unroll
		for( int k = 0; k < CA_UNROLL; k++ )
		{
			const int j = jb + k * JSTEP;
			u[ k ] = j < T ? K4[ (size_t)j * 8 + sub ] : make_uint4( 0, 0, 0, 0 );
		}

		ptx::pdl_wait();
		float qf[ 8 ];
#pragma unroll
		for( int e = 0; e < 8; e++ )
			qf[ e ] = __half2float( __float2half_rn( q[ (size_t)col * d + h * 64 + sub * 8 + e ] ) );

		// scores: a warp covers 4 key rows per load instruction (512 contiguous bytes), CA_UNROLL loads in flight per lane
		float lmax = -INFINITY;
		while( jb < T )
		{
#pragma unroll
			for( int k = 0; k < CA_UNROLL; k++ )
			{
				const int j = jb + k * JSTEP;
				const __half2* h2 = reinterpret_cast<const __half2*>( &u[ k ] );
				float s = 0.0f;
#pragma unroll
				for( int e = 0; e < 4; e++ )
				{
					const float2 f = __half22float2( h2[ e ] );
					s += f.x * qf[ e * 2 ] + f.y * qf[ e * 2 + 1 ];
				}
				s += __shfl_xor_sync( 0xffffffffu, s, 1 );
				s += __shfl_xor_sync( 0xffffffffu, s, 2 );
				s += __shfl_xor_sync( 0xffffffffu, s, 4 );
				if( j < T )
				{
					if( sub == 0 ) sp[ j ] = s;
					lmax = fmaxf( lmax, s );
				}
			}
			jb += JSTEP * CA_UNROLL;
			if( jb < T )
			{
#pragma unroll
				for( int k = 0; k < CA_UNROLL; k++ )
				{
					const int j = jb + k * JSTEP;
					u[ k ] = j < T ? K4[ (size_t)j * 8 + sub ] : make_uint4( 0, 0, 0, 0 );
				}
			}
		}
		lmax = warpMaxD( lmax );
		if( lane == 0 ) sred[ warp ] = lmax;
		__syncthreads();
		float mx = sred[ 0 ];
		for( int w = 1; w < CA_THREADS / 32; w++ ) mx = fmaxf( mx, sred[ w ] );
		__syncthreads();
		float lsum = 0.0f;
		for( int j = tid; j < T; j += CA_THREADS )
		{
			const float e = expF16Table( sp[ j ] - mx );
			sp[ j ] = e;
			lsum += e;
		}
		lsum = warpSumD( lsum );
		if( lane == 0 ) sred[ warp ] = lsum;
		__syncthreads();
		float tot = 0.0f;
		for( int w = 0; w < CA_THREADS / 32; w++ ) tot += sred[ w ];
		const float inv = 1.0f / tot;
		for( int j = tid; j < T; j += CA_THREADS ) sp[ j ] *= inv;   // normalised P (ggml.c:5085-5090)
		cp_async_wait_all();
		__syncthreads();

		const __half* sv = reinterpret_cast<const __half*>( ca_smem );
		if( refThreads > 0 )
		{
			// reference arithmetic (see pvChainF16): one f16 chain per (reference thread, output dim), V read from shared memory
			float* sof = &so[ 0 ][ 0 ][ 0 ];   // reused as [refThreads][64]
			const int dc = ( T + refThreads - 1 ) / refThreads;
			for( int idx = tid; idx < refThreads * 64; idx += CA_THREADS )
			{
				const int part = idx >> 6, e = idx & 63;
				const int j0 = min( part * dc, T ), j1 = min( ( part + 1 ) * dc, T );
				sof[ idx ] = pvChainF16( sp, sv + e, 64, j0, j1 );
			}
			__syncthreads();
			if( tid < 64 )
			{
				float acc = sof[ tid ];
				for( int k = 1; k < refThreads; k++ ) acc += sof[ k * 64 + tid ];
				out[ (size_t)col * d + h * 64 + tid ] = __float2half_rn( acc );
			}
			return;
		}

		// exact mode: f32 accumulation
		float o[ 8 ] = { 0, 0, 0, 0, 0, 0, 0, 0 };
		const uint4* sv4 = reinterpret_cast<const uint4*>( ca_smem );
		for( int j = warp * 4 + rgrp; j < T; j += ( CA_THREADS / 32 ) * 4 )
		{
			const float p = sp[ j ];
			const uint4 u = sv4[ (size_t)j * 8 + sub ];
			const __half2* h2 = reinterpret_cast<const __half2*>( &u );
#pragma unroll
			for( int e = 0; e < 4; e++ )
			{
				const float2 f = __half22float2( h2[ e ] );
				o[ e * 2 ] += p * f.x;
				o[ e * 2 + 1 ] += p * f.y;
			}
		}
#pragma unroll
		for( int e = 0; e < 8; e++ ) so[ warp ][ rgrp ][ sub * 8 + e ] = o[ e ];
		__syncthreads();
		if( tid < 64 )
		{
			float acc = 0.0f;
			for( int w = 0; w < CA_THREADS / 32; w++ )
#pragma unroll
				for( int r = 0; r < 4; r++ ) acc += so[ w ][ r ][ tid ];
			out[ (size_t)col * d + h * 64 + tid ] = __float2half_rn( acc );
		}
	}
	static PerDeviceMax g_crossAttr;
	static cudaError_t crossPrepare( int T )
	{
		return g_crossAttr.raise( (size_t)T * 128, []( size_t n ) { return cudaFuncSetAttribute( cross_attn_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)n ); } );
	}
	cudaError_t crossAttnDecode( const float* q, const __half* kMem, const __half* vMem, __half* out, int B, int N, int H, int d, int T, int refThreads, cudaStream_t s )
	{
		if( T > CA_MAXT || refThreads < 0 || refThreads > PV_MAX_THREADS ) return cudaErrorInvalidValue;
		cudaError_t e = crossPrepare( T );
		if( e != cudaSuccess ) return e;
		dim3 grid( B * H, N );
		return launchPdl( cross_attn_decode_kernel, grid, dim3( CA_THREADS ), (size_t)T * 128, s, q, kMem, vMem, out, N, H, d, T, refThreads );
	}

	// ---------------------------------------------------------------------------------------------------------------
	// softmax over the vocabulary + greedy sampling with the timestamp rules, one CTA per chunk.
	// Oracle: ggml_soft_max (ggml.c:5026-5095) then whisper_sample_best (whisper.cpp:1875-1964):
	//   max_tx = max p(text); if is_initial: p(ts > beg+100) = -inf; sum_ts, max_ts/tid over timestamp tokens;
	//   if sum_ts > max_tx or force_timestamp: p(text) = -inf;  top-4 by p; skip sot/solm/not among the first three.
	constexpr int SM_THREADS = 1024;

	// arg-max candidate; `tie`: another token has exactly the same probability (then the reference's pick depends on how
	// std::partial_sort happens to arrange equal keys, see emulatePartialSort)
	struct Top1 { float v; int i; int tie; };
	__device__ __forceinline__ Top1 better( Top1 a, Top1 b )
	{
		Top1 r = ( b.v > a.v || ( b.v == a.v && b.i < a.i ) ) ? b : a;
		if( a.v == b.v && a.i != b.i && a.v > -INFINITY ) r.tie = 1;
		return r;
	}

	// The reference ranks candidates with std::partial_sort( top 4 ) on (probability, id) pairs compared by probability only
	// (whisper.cpp:1932-1941).  With distinct probabilities that is "the largest"; with EXACT ties — typically a whole range of
	// probabilities that underflowed to 0 in the f16-table softmax — the winner is whatever libstdc++'s heap-select + sort-heap leaves
	// first.  This is that algorithm, statement by statement (bits/stl_heap.h: __push_heap, __adjust_heap, __make_heap, __sort_heap;
	// bits/stl_algo.h: __heap_select), run by ONE thread over the row: slow (~1 ms) and only entered when a tie decides the token.
	struct ProbId { double v; int i; };
	__device__ void heapPush( ProbId* f, int hole, int top, ProbId val )
	{
		int parent = ( hole - 1 ) / 2;
		while( hole > top && f[ parent ].v > val.v )
		{
			f[ hole ] = f[ parent ];
			hole = parent;
			parent = ( hole - 1 ) / 2;
		}
		f[ hole ] = val;
	}
	__device__ void heapAdjust( ProbId* f, int hole, int len, ProbId val )
	{
		const int top = hole;
		int second = hole;
		while( second < ( len - 1 ) / 2 )
		{
			second = 2 * ( second + 1 );
			if( f[ second ].v > f[ second - 1 ].v ) second--;
			f[ hole ] = f[ second ];
			hole = second;
		}
		if( ( len & 1 ) == 0 && second == ( len - 2 ) / 2 )
		{
			second = 2 * ( second + 1 );
			f[ hole ] = f[ second - 1 ];
			hole = second - 1;
		}
		heapPush( f, hole, top, val );
	}
	__device__ __noinline__ Top1 emulatePartialSort( const float* probs, int nv, int beg, bool maskText, bool isInitial, int sot, int solm, int nott )
	{
		auto val = [ & ]( int i, float p ) -> double {
			if( ( maskText && i < beg ) || ( isInitial && i >= beg + 101 ) ) return -(double)INFINITY;
			return (double)p;
		};
		ProbId f[ 4 ];
		for( int k = 0; k < 4; k++ ) f[ k ] = ProbId{ val( k, __ldcg( probs + k ) ), k };
		// __make_heap( first, first + 4 )
		for( int parent = ( 4 - 2 ) / 2;; parent-- )
		{
			const ProbId v = f[ parent ];
			heapAdjust( f, parent, 4, v );
			if( parent == 0 ) break;
		}
		// __heap_select: every later element that compares before the root replaces it
		for( int i = 4; i < nv; i++ )
		{
			const double v = val( i, __ldcg( probs + i ) );
			if( v > f[ 0 ].v ) heapAdjust( f, 0, 4, ProbId{ v, i } );
		}
		// __sort_heap
		for( int last = 3; last >= 1; last-- )
		{
			const ProbId v = f[ last ];
			f[ last ] = f[ 0 ];
			heapAdjust( f, 0, last, v );
		}
		int res = 0;
		while( ( f[ res ].i == sot || f[ res ].i == solm || f[ res ].i == nott ) && res < 3 ) res++;
		return Top1{ (float)f[ res ].v, f[ res ].i, 0 };
	}

	// The same result with the whole CTA at work.  std::__heap_select only ever moves an element into the heap that compares
	// strictly above the heap's root, and the root — the 4th largest value of the prefix processed so far — never decreases.  So
	// after thread 0 has run the exact algorithm over the first TIE_PRE elements, every later element that is not above the root
	// reached there (T1) can be dropped without changing a single heap operation: all warps filter their contiguous region of the row
	// against T1 (coalesced loads, ballot, survivors appended in index order to `scratch`), and thread 0 replays the exact algorithm
	// over the survivors only.  A row whose top probability is shared costs ~10 us this way instead of ~110 us on one thread
	// (degenerate rows tie at every step: that was 11 % of a decoder step).  Called by every thread of the CTA; the result is valid
	// in thread 0.  scratch: nv + 32 * warps ints.
	constexpr int TIE_PRE = 1024;
	struct TieShared
	{
		float t1;
		int tie, maskText;
		int cnt[ 32 ];
	};
	__device__ Top1 emulatePartialSortCta( const float* probs, int nv, int beg, bool maskText, bool isInitial, int sot, int solm, int nott, int* scratch, TieShared& sh, int nThreads )
	{
		auto val = [ & ]( int i, float p ) -> double {
			if( ( maskText && i < beg ) || ( isInitial && i >= beg + 101 ) ) return -(double)INFINITY;
			return (double)p;
		};
		const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nWarps = nThreads >> 5;
		const int pre = nv < TIE_PRE ? nv : TIE_PRE;
		ProbId f[ 4 ];
		if( tid == 0 )
		{
			for( int k = 0; k < 4; k++ ) f[ k ] = ProbId{ val( k, __ldcg( probs + k ) ), k };
			for( int parent = ( 4 - 2 ) / 2;; parent-- )
			{
				const ProbId v = f[ parent ];
				heapAdjust( f, parent, 4, v );
				if( parent == 0 ) break;
			}
			for( int i = 4; i < pre; i++ )
			{
				const double v = val( i, __ldcg( probs + i ) );
				if( v > f[ 0 ].v ) heapAdjust( f, 0, 4, ProbId{ v, i } );
			}
			sh.t1 = (float)f[ 0 ].v;    // (the values are floats or -inf: exact)
		}
		__syncthreads();
		const float t1 = sh.t1;
		const int region = ( ( ( nv - pre + nWarps - 1 ) / nWarps + 31 ) / 32 ) * 32;
		{
			const int i0 = pre + warp * region, i1 = min( nv, i0 + region );
			int cnt = 0;
			for( int base = i0; base < i1; base += 512 )
			{
				float v[ 16 ];   // sixteen independent loads in flight per lane: the filter is bound by L2 latency, not by work
#pragma unroll
				for( int u = 0; u < 16; u++ )
				{
					const int i = base + u * 32 + lane;
					v[ u ] = i < i1 ? (float)val( i, __ldcg( probs + i ) ) : -INFINITY;
				}
#pragma unroll
				for( int u = 0; u < 16; u++ )
				{
					const bool keep = v[ u ] > t1;
					const unsigned m = __ballot_sync( 0xffffffffu, keep );
					if( keep ) scratch[ warp * region + cnt + __popc( m & ( ( 1u << lane ) - 1u ) ) ] = base + u * 32 + lane;
					cnt += __popc( m );
				}
			}
			if( lane == 0 ) sh.cnt[ warp ] = cnt;
		}
		__syncthreads();
		if( tid != 0 ) return Top1{ 0.0f, 0, 0 };
		for( int w = 0; w < nWarps; w++ )
		{
			const int* list = scratch + w * region;
			const int n = sh.cnt[ w ];
			for( int k = 0; k < n; k++ )
			{
				const int i = list[ k ];
				const double v = val( i, __ldcg( probs + i ) );
				if( v > f[ 0 ].v ) heapAdjust( f, 0, 4, ProbId{ v, i } );
			}
		}
		for( int last = 3; last >= 1; last-- )
		{
			const ProbId v = f[ last ];
			f[ last ] = f[ 0 ];
			heapAdjust( f, 0, last, v );
		}
		int res = 0;
		while( ( f[ res ].i == sot || f[ res ].i == solm || f[ res ].i == nott ) && res < 3 ) res++;
		return Top1{ (float)f[ res ].v, f[ res ].i, 0 };
	}

	__device__ Top1 blockArgmax( Top1 x, Top1* sbuf )
	{
		for( int o = 16; o > 0; o >>= 1 )
		{
			Top1 y;
			y.v = __shfl_xor_sync( 0xffffffffu, x.v, o );
			y.i = __shfl_xor_sync( 0xffffffffu, x.i, o );
			y.tie = __shfl_xor_sync( 0xffffffffu, x.tie, o );
			x = better( x, y );
		}
		__syncthreads();
		if( ( threadIdx.x & 31 ) == 0 ) sbuf[ threadIdx.x >> 5 ] = x;
		__syncthreads();
		Top1 r = sbuf[ 0 ];
		for( int w = 1; w < SM_THREADS / 32; w++ ) r = better( r, sbuf[ w ] );
		return r;
	}
	__device__ float blockSumF( float v, float* sbuf )
	{
		v = warpSumD( v );
		__syncthreads();
		if( ( threadIdx.x & 31 ) == 0 ) sbuf[ threadIdx.x >> 5 ] = v;
		__syncthreads();
		float r = 0.0f;
		for( int w = 0; w < SM_THREADS / 32; w++ ) r += sbuf[ w ];
		return r;
	}
	__device__ float blockMaxF( float v, float* sbuf )
	{
		v = warpMaxD( v );
		__syncthreads();
		if( ( threadIdx.x & 31 ) == 0 ) sbuf[ threadIdx.x >> 5 ] = v;
		__syncthreads();
		float r = sbuf[ 0 ];
		for( int w = 1; w < SM_THREADS / 32; w++ ) r = fmaxf( r, sbuf[ w ] );
		return r;
	}
	__device__ double blockSumD( double v, double* sbuf )
	{
		for( int o = 16; o > 0; o >>= 1 ) v += __shfl_xor_sync( 0xffffffffu, v, o );
		__syncthreads();
		if( ( threadIdx.x & 31 ) == 0 ) sbuf[ threadIdx.x >> 5 ] = v;
		__syncthreads();
		double r = 0.0;
		for( int w = 0; w < SM_THREADS / 32; w++ ) r += sbuf[ w ];
		return r;
	}

	// `pick` is the arg-max over the eligible, non-banned tokens (lowest id among equals).  If its probability is shared — by another
	// eligible token or by one of the banned ones — the reference's choice depends on std::partial_sort's handling of equal keys:
	// reproduce it exactly.  The whole probability row must be visible in global memory (it is: written before the last barrier).
	__device__ __forceinline__ bool needsTieEmulation( Top1 pick, const SampleArgs& a, const float* probsRow, bool maskText )
	{
		bool tie = pick.tie != 0;
		if( !tie && !maskText )
		{
			const int banned[ 3 ] = { a.tokenSot, a.tokenSolm, a.tokenNot };
			for( int k = 0; k < 3; k++ )
				if( banned[ k ] >= 0 && banned[ k ] < a.tokenBeg && __ldcg( probsRow + banned[ k ] ) == pick.v ) tie = true;
		}
		return tie;
	}
	// called by every thread of the CTA that finishes a row; `pick` and the result are thread 0's
	__device__ __forceinline__ Top1 resolveTies( Top1 pick, const SampleArgs& a, int b, bool maskText, bool isInitial, TieShared& sh )
	{
		const float* probsRow = a.probs + (size_t)b * a.nVocab;
		if( threadIdx.x == 0 ) sh.tie = needsTieEmulation( pick, a, probsRow, maskText ) ? 1 : 0;
		__syncthreads();
		if( !sh.tie ) return pick;
		if( !a.tieScratch )
		{
			if( threadIdx.x == 0 ) pick = emulatePartialSort( probsRow, a.nVocab, a.tokenBeg, maskText, isInitial, a.tokenSot, a.tokenSolm, a.tokenNot );
			return pick;
		}
		const Top1 r = emulatePartialSortCta( probsRow, a.nVocab, a.tokenBeg, maskText, isInitial, a.tokenSot, a.tokenSolm, a.tokenNot,
			a.tieScratch + (size_t)b * ( a.nVocab + 32 * ( SM_THREADS / 32 ) ), sh, SM_THREADS );
		return threadIdx.x == 0 ? r : pick;
	}

	__global__ void __launch_bounds__( SM_THREADS )
		sample_kernel( SampleArgs a )
	{
		__shared__ float sf[ SM_THREADS / 32 ];
		__shared__ double sd[ SM_THREADS / 32 ];
		__shared__ Top1 st[ SM_THREADS / 32 ];
		ptx::pdl_launch_dependents();
		ptx::pdl_wait();
		const int b = blockIdx.x;
		const int tid = threadIdx.x;
		const int nv = a.nVocab;
		const float* lg = a.logits + (size_t)b * nv;
		float* gpr = a.probs + (size_t)b * nv;
		// the row lives in shared memory when it fits (207 KB for 51865 tokens): seven passes over L2 cost 65 us, over smem ~15 us
		extern __shared__ float sampleRow[];
		float* pr = a.rowInSmem ? sampleRow : gpr;
		const bool forceTs = a.dForceTs && a.dForceTs[ 0 ] != 0;
		const bool isInitial = a.dForceTs && a.dForceTs[ 1 ] != 0;

		float mx = -INFINITY;
		for( int i = tid; i < nv; i += SM_THREADS )
		{
			const float v = __ldcg( lg + i );
			pr[ i ] = v;
			mx = fmaxf( mx, v );
		}
		mx = blockMaxF( mx, sf );
		double dsum = 0.0;
		for( int i = tid; i < nv; i += SM_THREADS )
		{
			const float e = expF16Table( pr[ i ] - mx );
			pr[ i ] = e;
			dsum += (double)e;
		}
		dsum = blockSumD( dsum, sd );
		const float inv = (float)( 1.0 / dsum );
		// timestamp bookkeeping on the normalised probabilities
		const int beg = a.tokenBeg;
		const int tsEnd = isInitial ? min( beg + 101, nv ) : nv;   // initial timestamp <= 1 s (whisper.cpp:1902-1911)
		float maxTx = -1.0f;
		double sumTs = 0.0;
		// The reference ranks the top 4 and skips sot / solm / not at most 3 times (whisper.cpp:1945-1953); those are the only three
		// banned ids, so the pick is simply the arg-max over every other eligible token.  Eligibility depends on the timestamp test
		// below, hence two candidates from the same pass: the best text token and the best timestamp token.
		Top1 bestTs = { -INFINITY, 0x7fffffff, 0 };
		Top1 bestTx = { -INFINITY, 0x7fffffff, 0 };
		for( int i = tid; i < nv; i += SM_THREADS )
		{
			const float p = pr[ i ] * inv;
			if( a.rowInSmem ) gpr[ i ] = p; else pr[ i ] = p;
			if( i < beg )
			{
				maxTx = fmaxf( maxTx, p );
				if( i != a.tokenSot && i != a.tokenSolm && i != a.tokenNot ) bestTx = better( bestTx, Top1{ p, i, 0 } );
			}
			else if( i < tsEnd )
			{
				sumTs += (double)p;
				bestTs = better( bestTs, Top1{ p, i, 0 } );
			}
		}
		maxTx = blockMaxF( maxTx, sf );
		sumTs = blockSumD( sumTs, sd );
		bestTs = blockArgmax( bestTs, st );
		bestTx = blockArgmax( bestTx, st );
		const bool maskText = ( sumTs > (double)maxTx ) || forceTs;
		__shared__ TieShared tieSh;
		Top1 pick = bestTs;
		if( !maskText ) pick = better( bestTx, bestTs );
		if( pick.i == 0x7fffffff ) pick = Top1{ 0.0f, 0, 1 };
		__syncthreads();     // the whole probability row is in global memory
		pick = resolveTies( pick, a, b, maskText, isInitial, tieSh );
		if( tid == 0 )
		{
			TokenData td;
			td.id = pick.i;
			td.tid = bestTs.i == 0x7fffffff ? 0 : bestTs.i;
			td.p = pick.v;
			td.pt = (float)( (double)bestTs.v / ( sumTs + 1e-10 ) );
			td.ptsum = (float)sumTs;
			a.out[ b ] = td;
			if( a.nextTokens ) a.nextTokens[ b ] = td.id;
			if( a.history && a.dStep && *a.dStep < a.histCap ) a.history[ (size_t)b * a.histCap + *a.dStep ] = td.id;
		}
	}

	// The same sampler with one vocabulary row spread over a cluster of SC_CL CTAs: every CTA keeps its slice of the row in shared
	// memory, the three row-wide reductions go through distributed shared memory (each CTA publishes its partial, cluster barrier,
	// everybody combines the SC_CL partials in rank order).  8 rows on 8 SMs left 140 SMs idle for 37 us per token step.
	struct ClusterPart
	{
		float mx, maxTx;
		double dsum, sumTs;
		Top1 bestTs, bestTx;
	};
	template<int SC_CL>
	__global__ void __launch_bounds__( SM_THREADS )
		sample_cluster_kernel( SampleArgs a )
	{
		namespace cg = cooperative_groups;
		cg::cluster_group cluster = cg::this_cluster();
		const unsigned rank = cluster.block_rank();
		__shared__ ClusterPart part;
		__shared__ float sf[ SM_THREADS / 32 ];
		__shared__ double sd[ SM_THREADS / 32 ];
		__shared__ Top1 st[ SM_THREADS / 32 ];
		extern __shared__ float sampleRow[];
		const int b = blockIdx.y;
		const int tid = threadIdx.x;
		const int nv = a.nVocab;
		const int per = ( nv + SC_CL - 1 ) / SC_CL;
		const int i0 = min( (int)rank * per, nv ), i1 = min( i0 + per, nv );
		const int n = i1 - i0;
		const float* lg = a.logits + (size_t)b * nv + i0;
		float* gpr = a.probs + (size_t)b * nv + i0;
		float* pr = sampleRow;
		const bool forceTs = a.dForceTs && a.dForceTs[ 0 ] != 0;
		const bool isInitial = a.dForceTs && a.dForceTs[ 1 ] != 0;

		float mx = -INFINITY;
		for( int i = tid; i < n; i += SM_THREADS )
		{
			const float v = __ldcg( lg + i );
			pr[ i ] = v;
			mx = fmaxf( mx, v );
		}
		mx = blockMaxF( mx, sf );
		if( tid == 0 ) part.mx = mx;
		cluster.sync();
		for( unsigned r = 0; r < SC_CL; r++ ) mx = fmaxf( mx, cluster.map_shared_rank( &part, r )->mx );

		double dsum = 0.0;
		for( int i = tid; i < n; i += SM_THREADS )
		{
			const float e = expF16Table( pr[ i ] - mx );
			pr[ i ] = e;
			dsum += (double)e;
		}
		dsum = blockSumD( dsum, sd );
		if( tid == 0 ) part.dsum = dsum;
		cluster.sync();
		double total = 0.0;   // the terms are f16 values: this double sum is exact, so the split does not change it
		for( unsigned r = 0; r < SC_CL; r++ ) total += cluster.map_shared_rank( &part, r )->dsum;
		const float inv = (float)( 1.0 / total );

		const int beg = a.tokenBeg;
		const int tsEnd = isInitial ? min( beg + 101, nv ) : nv;
		float maxTx = -1.0f;
		double sumTs = 0.0;
		Top1 bestTs = { -INFINITY, 0x7fffffff, 0 };
		Top1 bestTx = { -INFINITY, 0x7fffffff, 0 };
		for( int i = tid; i < n; i += SM_THREADS )
		{
			const float p = pr[ i ] * inv;
			gpr[ i ] = p;
			const int gi = i0 + i;
			if( gi < beg )
			{
				maxTx = fmaxf( maxTx, p );
				if( gi != a.tokenSot && gi != a.tokenSolm && gi != a.tokenNot ) bestTx = better( bestTx, Top1{ p, gi, 0 } );
			}
			else if( gi < tsEnd )
			{
				sumTs += (double)p;
				bestTs = better( bestTs, Top1{ p, gi, 0 } );
			}
		}
		maxTx = blockMaxF( maxTx, sf );
		sumTs = blockSumD( sumTs, sd );
		bestTs = blockArgmax( bestTs, st );
		bestTx = blockArgmax( bestTx, st );
		if( tid == 0 )
		{
			part.maxTx = maxTx; part.sumTs = sumTs; part.bestTs = bestTs; part.bestTx = bestTx;
		}
		cluster.sync();
		__shared__ TieShared tieSh;
		if( rank == 0 )
		{
			// the row is finished by the first CTA of its cluster: thread 0 combines the partials, all threads help if the top
			// probability is shared (resolveTies)
			Top1 pick = bestTs;
			if( tid == 0 )
			{
				for( unsigned r = 1; r < SC_CL; r++ )
				{
					const ClusterPart* q = cluster.map_shared_rank( &part, r );
					maxTx = fmaxf( maxTx, q->maxTx );
					sumTs += q->sumTs;
					bestTs = better( bestTs, q->bestTs );
					bestTx = better( bestTx, q->bestTx );
				}
				const bool mt = ( sumTs > (double)maxTx ) || forceTs;
				pick = bestTs;
				if( !mt ) pick = better( bestTx, bestTs );
				if( pick.i == 0x7fffffff ) pick = Top1{ 0.0f, 0, 1 };
				tieSh.maskText = mt ? 1 : 0;
			}
			__syncthreads();
			const bool maskText = tieSh.maskText != 0;
			pick = resolveTies( pick, a, b, maskText, isInitial, tieSh );
		if( tid == 0 )
		{
			TokenData td;
			td.id = pick.i;
			td.tid = bestTs.i == 0x7fffffff ? 0 : bestTs.i;
			td.p = pick.v;
			td.pt = (float)( (double)bestTs.v / ( sumTs + 1e-10 ) );
			td.ptsum = (float)sumTs;
			a.out[ b ] = td;
			if( a.nextTokens ) a.nextTokens[ b ] = td.id;
			if( a.history && a.dStep && *a.dStep < a.histCap ) a.history[ (size_t)b * a.histCap + *a.dStep ] = td.id;
		}
		}
		cluster.sync();   // nobody leaves while a peer may still read its partials
	}

	__global__ void __launch_bounds__( SM_THREADS )
		softmax_rows_kernel( const float* __restrict__ logits, float* __restrict__ probs, int n )
	{
		__shared__ float sf[ SM_THREADS / 32 ];
		__shared__ double sd[ SM_THREADS / 32 ];
		const float* lg = logits + (size_t)blockIdx.x * n;
		float* pr = probs + (size_t)blockIdx.x * n;
		const int tid = threadIdx.x;
		float mx = -INFINITY;
		for( int i = tid; i < n; i += SM_THREADS ) mx = fmaxf( mx, lg[ i ] );
		mx = blockMaxF( mx, sf );
		double dsum = 0.0;
		for( int i = tid; i < n; i += SM_THREADS )
		{
			const float e = expF16Table( lg[ i ] - mx );
			pr[ i ] = e;
			dsum += (double)e;
		}
		dsum = blockSumD( dsum, sd );
		const float inv = (float)( 1.0 / dsum );
		for( int i = tid; i < n; i += SM_THREADS ) pr[ i ] *= inv;
	}
	cudaError_t softmaxRows( const float* logits, float* probs, int rows, int n, cudaStream_t s )
	{
		softmax_rows_kernel<<<rows, SM_THREADS, 0, s>>>( logits, probs, n );
		return cudaGetLastError();
	}

	__global__ void advance_kernel( int* dNPast, int N, int* dForceTs, int* dStep )
	{
		ptx::pdl_launch_dependents();
		ptx::pdl_wait();
		if( dNPast ) *dNPast += N;
		if( dForceTs ) { dForceTs[ 0 ] = 0; dForceTs[ 1 ] = 0; }
		if( dStep ) *dStep += 1;
	}

	cudaError_t sampleGreedy( const SampleArgs& a, cudaStream_t s )
	{
		SampleArgs sa = a;
		// cluster size: 4 by default (measured 157.0 -> 155.4 ms per 100 tokens against the single-CTA kernel); WSP_SAMPLER_CLUSTER=0|4|8 for A/B runs
		static const int clusterSize = []() { const char* e = getenv( "WSP_SAMPLER_CLUSTER" ); return e ? atoi( e ) : 4; }();
		if( ( clusterSize == 4 || clusterSize == 8 ) && a.nVocab >= 4096 )
		{
			const int cl = clusterSize;
			const size_t sliceBytes = (size_t)( ( a.nVocab + cl - 1 ) / cl ) * sizeof( float );
			static PerDeviceMax sliceAttr;
			{
				const int nv = a.nVocab;
				cudaError_t ea = sliceAttr.raise( (size_t)( ( nv + 3 ) / 4 ) * sizeof( float ), [ nv ]( size_t n4 ) {
					cudaError_t e4 = cudaFuncSetAttribute( sample_cluster_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)n4 );
					if( e4 != cudaSuccess ) return e4;
					return cudaFuncSetAttribute( sample_cluster_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)( (size_t)( ( nv + 7 ) / 8 ) * sizeof( float ) ) );
				} );
				if( ea != cudaSuccess ) return ea;
			}
			cudaLaunchConfig_t cfg{};
			cfg.gridDim = dim3( cl, a.B );
			cfg.blockDim = dim3( SM_THREADS );
			cfg.dynamicSmemBytes = sliceBytes;
			cfg.stream = s;
			cudaLaunchAttribute at[ 1 ];
			at[ 0 ].id = cudaLaunchAttributeClusterDimension;
			at[ 0 ].val.clusterDim.x = cl; at[ 0 ].val.clusterDim.y = 1; at[ 0 ].val.clusterDim.z = 1;
			cfg.attrs = at;
			cfg.numAttrs = 1;
			cudaError_t e = cl == 4 ? cudaLaunchKernelEx( &cfg, sample_cluster_kernel<4>, sa ) : cudaLaunchKernelEx( &cfg, sample_cluster_kernel<8>, sa );
			if( e != cudaSuccess ) return e;
			return launchPdl( advance_kernel, dim3( 1 ), dim3( 1 ), 0, s, a.dNPast, a.N, const_cast<int*>( a.dForceTs ), a.dStep );
		}
		const size_t rowBytes = (size_t)a.nVocab * sizeof( float );
		static const bool forceGlobal = getenv( "WSP_SAMPLER_GLOBAL" ) != nullptr;   // A/B switch for measurements
		sa.rowInSmem = ( rowBytes <= 220 * 1024 && !forceGlobal ) ? 1 : 0;
		static PerDeviceMax rowAttr;
		if( sa.rowInSmem )
		{
			cudaError_t ea = rowAttr.raise( rowBytes, []( size_t n ) { return cudaFuncSetAttribute( sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)n ); } );
			if( ea != cudaSuccess ) return ea;
		}
		cudaError_t e = launchPdl( sample_kernel, dim3( a.B ), dim3( SM_THREADS ), sa.rowInSmem ? rowBytes : 0, s, sa );
		if( e != cudaSuccess ) return e;
		// step-global state is advanced by a separate 1-thread kernel so that no CTA of the sampler races with it
		return launchPdl( advance_kernel, dim3( 1 ), dim3( 1 ), 0, s, a.dNPast, a.N, const_cast<int*>( a.dForceTs ), a.dStep );
	}
}
