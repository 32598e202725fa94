// tcgen05 + TMA GEMM for sm_100a — see gemm_tc.cuh for the design notes.
#include "gemm_tc.cuh"
#include "per_device.h"
#include <stdio.h>
#include <mutex>

namespace gemm
{
	// ---------------------------------------------------------------------------------------------------------------
	// epilogues: one thread owns output row `m` and 32 consecutive columns [n0, n0+32)
	__device__ __forceinline__ void store_f16x32( __half* dst, const float* v )
	{
		uint4* p = reinterpret_cast<uint4*>( dst );
#pragma unroll
		for( int i = 0; i < 4; i++ )
		{
			__half2 h0 = __floats2half2_rn( v[ i * 8 + 0 ], v[ i * 8 + 1 ] );
			__half2 h1 = __floats2half2_rn( v[ i * 8 + 2 ], v[ i * 8 + 3 ] );
			__half2 h2 = __floats2half2_rn( v[ i * 8 + 4 ], v[ i * 8 + 5 ] );
			__half2 h3 = __floats2half2_rn( v[ i * 8 + 6 ], v[ i * 8 + 7 ] );
			uint4 u;
			u.x = *reinterpret_cast<uint32_t*>( &h0 );
			u.y = *reinterpret_cast<uint32_t*>( &h1 );
			u.z = *reinterpret_cast<uint32_t*>( &h2 );
			u.w = *reinterpret_cast<uint32_t*>( &h3 );
			p[ i ] = u;
		}
	}

	template<int MODE>
	__device__ __forceinline__ void epilogue( const EpiParams& ep, int m, int n0, float* v )
	{
		if( n0 >= ep.N )
			return;
		if constexpr( MODE == EPI_F32 )
		{
			if( m >= ep.M ) return;
			float* dst = ep.out_f32 + (size_t)m * ep.ld + n0;
			if( n0 + 32 <= ep.N && ( ep.ld & 3 ) == 0 )
			{
#pragma unroll
				for( int i = 0; i < 8; i++ )
				{
					float4 o;
					o.x = v[ i * 4 + 0 ]; o.y = v[ i * 4 + 1 ]; o.z = v[ i * 4 + 2 ]; o.w = v[ i * 4 + 3 ];
					if( ep.bias )
					{
						const float4 b = *reinterpret_cast<const float4*>( ep.bias + n0 + i * 4 );
						o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
					}
					reinterpret_cast<float4*>( dst )[ i ] = o;
				}
			}
			else
			{
				for( int i = 0; i < 32 && n0 + i < ep.N; i++ )
					dst[ i ] = v[ i ] + ( ep.bias ? ep.bias[ n0 + i ] : 0.0f );
			}
		}
		else if constexpr( MODE == EPI_CONV1 )
		{
			const int b = m / ep.rows_per_chunk;
			const int t = m - b * ep.rows_per_chunk;
			if( b >= ep.nchunks || t >= ep.valid_per_chunk ) return;
#pragma unroll
			for( int i = 0; i < 32; i++ )
				v[ i ] = ptx::gelu_f16_semantics( v[ i ] + ep.bias[ n0 + i ] );
			store_f16x32( ep.out_a + (size_t)( m + 1 ) * ep.ld + n0, v );
		}
		else if constexpr( MODE == EPI_CONV2 )
		{
			const int b = m / ep.rows_per_chunk;
			const int j = m - b * ep.rows_per_chunk;
			if( b >= ep.nchunks || j >= ep.valid_per_chunk ) return;
			float* dst = ep.out_f32 + ( (size_t)b * ep.T + j ) * ep.ld + n0;
			const float* pe = ep.pos + (size_t)j * ep.d + n0;
#pragma unroll
			for( int i = 0; i < 8; i++ )
			{
				const float4 p = *reinterpret_cast<const float4*>( pe + i * 4 );
				float4 o;
				o.x = ptx::gelu_f16_semantics( v[ i * 4 + 0 ] + ep.bias[ n0 + i * 4 + 0 ] ) + p.x;
				o.y = ptx::gelu_f16_semantics( v[ i * 4 + 1 ] + ep.bias[ n0 + i * 4 + 1 ] ) + p.y;
				o.z = ptx::gelu_f16_semantics( v[ i * 4 + 2 ] + ep.bias[ n0 + i * 4 + 2 ] ) + p.z;
				o.w = ptx::gelu_f16_semantics( v[ i * 4 + 3 ] + ep.bias[ n0 + i * 4 + 3 ] ) + p.w;
				reinterpret_cast<float4*>( dst )[ i ] = o;
			}
		}
		else if constexpr( MODE == EPI_QKV )
		{
			if( m >= ep.M ) return;
			const int b = m / ep.T;
			const int t = m - b * ep.T;
			const int which = n0 / ep.d;
			const int nn = n0 - which * ep.d;
			const int h = nn >> 6;
			const int e = nn & 63;
#pragma unroll
			for( int i = 0; i < 32; i++ )
				v[ i ] += ep.bias[ n0 + i ];
			if( which < 2 )
			{
				__half* base = which == 0 ? ep.out_a : ep.out_b;
				store_f16x32( base + ( ( (size_t)b * ep.H + h ) * ep.T + t ) * 64 + e, v );
			}
			else
			{
				__half* dst = ep.out_c + ( ( (size_t)b * ep.H + h ) * 64 + e ) * ep.Tp + t;
#pragma unroll
				for( int i = 0; i < 32; i++ )
					dst[ (size_t)i * ep.Tp ] = __float2half_rn( v[ i ] );
			}
		}
		else if constexpr( MODE == EPI_BIAS_RESID )
		{
			if( m >= ep.M ) return;
			const size_t off = (size_t)m * ep.ld + n0;
#pragma unroll
			for( int i = 0; i < 8; i++ )
			{
				const float4 b = *reinterpret_cast<const float4*>( ep.bias + n0 + i * 4 );
				const float4 r = *reinterpret_cast<const float4*>( ep.resid + off + i * 4 );
				float4 o;
				o.x = v[ i * 4 + 0 ] + b.x + r.x;
				o.y = v[ i * 4 + 1 ] + b.y + r.y;
				o.z = v[ i * 4 + 2 ] + b.z + r.z;
				o.w = v[ i * 4 + 3 ] + b.w + r.w;
				*reinterpret_cast<float4*>( ep.out_f32 + off + i * 4 ) = o;
			}
		}
		else if constexpr( MODE == EPI_BIAS_GELU_F16 )
		{
			if( m >= ep.M ) return;
#pragma unroll
			for( int i = 0; i < 32; i++ )
				v[ i ] = ptx::gelu_f16_semantics( v[ i ] + ep.bias[ n0 + i ] );
			store_f16x32( ep.out_a + (size_t)m * ep.ld + n0, v );
		}
		else if constexpr( MODE == EPI_CROSSKV )
		{
			if( m >= ep.M ) return;
			const int b = m / ep.T;
			const int t = m - b * ep.T;
			const int d2 = 2 * ep.d;
			const int l = n0 / d2;
			const int r = n0 - l * d2;
			const bool isV = r >= ep.d;
			const int nn = isV ? r - ep.d : r;
			const int h = nn >> 6;
			const int e = nn & 63;
			if( isV )
			{
#pragma unroll
				for( int i = 0; i < 32; i++ )
					v[ i ] += ep.bias[ n0 + i ];
			}
			else
			{
#pragma unroll
				for( int i = 0; i < 32; i++ )
					v[ i ] *= ep.scale;
			}
			__half* base = isV ? ep.out_b : ep.out_a;
			store_f16x32( base + ( ( ( (size_t)l * ep.nchunks + b ) * ep.H + h ) * ep.T + t ) * 64 + e, v );
		}
	}

	// ---------------------------------------------------------------------------------------------------------------
	// Coalesced epilogues.  tcgen05.ld hands every thread one ROW of the accumulator tile, so a direct store makes each warp
	// instruction touch 32 different rows (16 bytes each): half of every 32-byte sector is wasted and a 128-byte line needs 8
	// separate instructions (measured: the out-projection GEMM ran at 17 % tensor-pipe utilisation, epilogue-bound).  The
	// accumulator chunk (32 rows x 32 columns) is therefore transposed through a padded per-warp shared-memory tile so that a
	// thread owns one COLUMN: every load / store instruction of the warp then covers 32 consecutive elements of one output row.
	__device__ __forceinline__ void loadResid( const EpiParams& ep, int mBase, int n, float* rs )
	{
#pragma unroll
		for( int rr = 0; rr < 32; rr++ )
		{
			const int m = mBase + rr;
			rs[ rr ] = ( m < ep.M && n < ep.N ) ? ep.resid[ (size_t)m * ep.ld + n ] : 0.0f;
		}
	}
	template<int MODE>
	__device__ __forceinline__ void epilogueT( const EpiParams& ep, int mBase, int n, const float* w, const float* rs = nullptr )
	{
		if( n >= ep.N ) return;
		if constexpr( MODE == EPI_F32 )
		{
			const float b = ep.bias ? ep.bias[ n ] : 0.0f;
#pragma unroll
			for( int rr = 0; rr < 32; rr++ )
			{
				const int m = mBase + rr;
				if( m < ep.M ) ep.out_f32[ (size_t)m * ep.ld + n ] = w[ rr ] + b;
			}
		}
		else if constexpr( MODE == EPI_CONV1 )
		{
			const float b = ep.bias[ n ];
			int cb = mBase / ep.rows_per_chunk;
			int t = mBase - cb * ep.rows_per_chunk;
#pragma unroll
			for( int rr = 0; rr < 32; rr++ )
			{
				if( cb < ep.nchunks && t < ep.valid_per_chunk )
					ep.out_a[ (size_t)( mBase + rr + 1 ) * ep.ld + n ] = __float2half_rn( ptx::gelu_f16_semantics( w[ rr ] + b ) );
				if( ++t == ep.rows_per_chunk ) { t = 0; cb++; }
			}
		}
		else if constexpr( MODE == EPI_CONV2 )
		{
			const float b = ep.bias[ n ];
			const int cb0 = mBase / ep.rows_per_chunk;
			const int j0 = mBase - cb0 * ep.rows_per_chunk;
			// positional rows first (loads must not be serialised behind the stores: pos and out_f32 are both float*)
			float pe[ 32 ];
			{
				int j = j0;
#pragma unroll
				for( int rr = 0; rr < 32; rr++ )
				{
					pe[ rr ] = j < ep.valid_per_chunk ? ep.pos[ (size_t)j * ep.d + n ] : 0.0f;
					if( ++j == ep.rows_per_chunk ) j = 0;
				}
			}
			int cb = cb0, j = j0;
#pragma unroll
			for( int rr = 0; rr < 32; rr++ )
			{
				if( cb < ep.nchunks && j < ep.valid_per_chunk )
					ep.out_f32[ ( (size_t)cb * ep.T + j ) * ep.ld + n ] = ptx::gelu_f16_semantics( w[ rr ] + b ) + pe[ rr ];
				if( ++j == ep.rows_per_chunk ) { j = 0; cb++; }
			}
		}
		else if constexpr( MODE == EPI_QKV )
		{
			// Q and K only (the V^T part keeps the row-per-thread layout, which is the coalesced one for a transposed store)
			const int which = n / ep.d;
			const int nn = n - which * ep.d;
			const int h = nn >> 6, e = nn & 63;
			const float b = ep.bias[ n ];
			__half* base = which == 0 ? ep.out_a : ep.out_b;
			int cb = mBase / ep.T;
			int t = mBase - cb * ep.T;
#pragma unroll
			for( int rr = 0; rr < 32; rr++ )
			{
				if( mBase + rr < ep.M )
					base[ ( ( (size_t)cb * ep.H + h ) * ep.T + t ) * 64 + e ] = __float2half_rn( w[ rr ] + b );
				if( ++t == ep.T ) { t = 0; cb++; }
			}
		}
		else if constexpr( MODE == EPI_BIAS_RESID )
		{
			// rs = the residual column, fetched by the caller one 32-column chunk ahead (loadResid): the 32 loads of a chunk are a
			// full L2 round trip, and with only 4 epilogue warps that latency - not bandwidth - set the epilogue time (16 us per tile)
			const float b = ep.bias[ n ];
#pragma unroll
			for( int rr = 0; rr < 32; rr++ )
			{
				const int m = mBase + rr;
				if( m < ep.M ) ep.out_f32[ (size_t)m * ep.ld + n ] = w[ rr ] + b + rs[ rr ];
			}
		}
		else if constexpr( MODE == EPI_BIAS_GELU_F16 )
		{
			const float b = ep.bias[ n ];
#pragma unroll
			for( int rr = 0; rr < 32; rr++ )
			{
				const int m = mBase + rr;
				if( m < ep.M ) ep.out_a[ (size_t)m * ep.ld + n ] = __float2half_rn( ptx::gelu_f16_semantics( w[ rr ] + b ) );
			}
		}
		else if constexpr( MODE == EPI_CROSSKV )
		{
			const int d2 = 2 * ep.d;
			const int l = n / d2;
			const int r = n - l * d2;
			const bool isV = r >= ep.d;
			const int nn = isV ? r - ep.d : r;
			const int h = nn >> 6, e = nn & 63;
			const float b = isV ? ep.bias[ n ] : 0.0f;
			const float sc = isV ? 1.0f : ep.scale;
			__half* base = isV ? ep.out_b : ep.out_a;
			int cb = mBase / ep.T;
			int t = mBase - cb * ep.T;
#pragma unroll
			for( int rr = 0; rr < 32; rr++ )
			{
				if( mBase + rr < ep.M )
					base[ ( ( ( (size_t)l * ep.nchunks + cb ) * ep.H + h ) * ep.T + t ) * 64 + e ] = __float2half_rn( w[ rr ] * sc + b );
				if( ++t == ep.T ) { t = 0; cb++; }
			}
		}
	}

	// ---------------------------------------------------------------------------------------------------------------
	template<int BN, int STAGES>
	struct SmemLayout
	{
		static constexpr int A_BYTES = BM * BK * 2;
		static constexpr int B_BYTES = BN * BK * 2;
		static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
		static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
		static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;   // + barriers + alignment slack (the transpose tiles are static smem)
	};

	template<int BN, int STAGES, int MODE, int AMODE>
	__global__ void __launch_bounds__( 192, 1 )
		gemm_tc_kernel( const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapB,
			int M, int N, int K, int kTap, EpiParams ep )
	{
		using SL = SmemLayout<BN, STAGES>;
		extern __shared__ uint8_t smem_raw[];
		__shared__ float xposeTiles[ 4 ][ 32 * 33 ];   // one padded 32x32 transpose tile per epilogue warp (static: compiled to LDS/STS)
		uint8_t* smem = reinterpret_cast<uint8_t*>( ( reinterpret_cast<uintptr_t>( smem_raw ) + 1023 ) & ~(uintptr_t)1023 );
		uint64_t* bar_full = reinterpret_cast<uint64_t*>( smem + SL::BAR_OFFSET );
		uint64_t* bar_empty = bar_full + STAGES;
		uint64_t* bar_tfull = bar_empty + STAGES;
		uint64_t* bar_tempty = bar_tfull + 2;
		uint32_t* tmem_slot = reinterpret_cast<uint32_t*>( bar_tempty + 2 );

		const int warp = threadIdx.x >> 5;
		const int lane = threadIdx.x & 31;

		const int num_m = ( M + BM - 1 ) / BM;
		const int num_n = ( N + BN - 1 ) / BN;
		const int num_tiles = num_m * num_n;
		const int num_kb = ( K + BK - 1 ) / BK;
		constexpr uint32_t TMEM_COLS = 2 * BN;   // two accumulator stages (power of two: 256 or 512)

		if( warp == 0 && lane == 0 )
		{
			ptx::prefetch_tensormap( &mapA );
			ptx::prefetch_tensormap( &mapB );
			if( AMODE == A_CONV_S2 ) ptx::prefetch_tensormap( &mapA2 );
			for( int s = 0; s < STAGES; s++ )
			{
				ptx::mbar_init( &bar_full[ s ], 1 );
				ptx::mbar_init( &bar_empty[ s ], 1 );
			}
			for( int s = 0; s < 2; s++ )
			{
				ptx::mbar_init( &bar_tfull[ s ], 1 );
				ptx::mbar_init( &bar_tempty[ s ], 4 );   // one arrive per epilogue warp
			}
			ptx::fence_barrier_init();
		}
		if( warp == 1 )
		{
			ptx::tmem_alloc( tmem_slot, TMEM_COLS );
			ptx::tmem_relinquish();
		}
		ptx::tc_fence_before();
		__syncthreads();
		ptx::tc_fence_after();
		const uint32_t tmem_base = *tmem_slot;

		if( warp == 0 )
		{
			// ===== TMA producer =====
			if( lane == 0 )
			{
				int stage = 0;
				uint32_t phase = 0;
				for( int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x )
				{
					const int m0 = ( tile / num_n ) * BM;
					const int n0 = ( tile % num_n ) * BN;
					for( int kb = 0; kb < num_kb; kb++ )
					{
						ptx::mbar_wait( &bar_empty[ stage ], phase ^ 1 );
						ptx::mbar_expect_tx( &bar_full[ stage ], SL::STAGE_BYTES );
						uint8_t* sa = smem + stage * SL::STAGE_BYTES;
						uint8_t* sb = sa + SL::A_BYTES;
						if constexpr( AMODE == A_PLAIN )
							ptx::tma_load_2d( sa, &mapA, &bar_full[ stage ], kb * BK, m0 );
						else
						{
							const int tap = kb / kTap;
							const int c0 = ( kb - tap * kTap ) * BK;
							if constexpr( AMODE == A_CONV_S1 )
								ptx::tma_load_2d( sa, &mapA, &bar_full[ stage ], c0, m0 + tap );
							else
							{
								if( tap == 1 )
									ptx::tma_load_2d( sa, &mapA2, &bar_full[ stage ], c0, m0 );
								else
									ptx::tma_load_2d( sa, &mapA, &bar_full[ stage ], c0, m0 + ( tap >> 1 ) );
							}
						}
						ptx::tma_load_2d( sb, &mapB, &bar_full[ stage ], kb * BK, n0 );
						if( ++stage == STAGES ) { stage = 0; phase ^= 1; }
					}
				}
			}
			__syncwarp();
		}
		else if( warp == 1 )
		{
			// ===== MMA issuer =====
			if( lane == 0 )
			{
				constexpr uint32_t idesc = ptx::umma_idesc_f16( BM, BN );
				int stage = 0;
				uint32_t phase = 0;
				int as = 0;
				uint32_t aphase = 0;
				for( int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x )
				{
					ptx::mbar_wait( &bar_tempty[ as ], aphase ^ 1 );
					ptx::tc_fence_after();
					const uint32_t tmem_d = tmem_base + (uint32_t)( as * BN );
					for( int kb = 0; kb < num_kb; kb++ )
					{
						ptx::mbar_wait( &bar_full[ stage ], phase );
						ptx::tc_fence_after();
						const uint32_t sa = ptx::smem_u32( smem + stage * SL::STAGE_BYTES );
						const uint64_t da = ptx::umma_desc_sw128( sa );
						const uint64_t db = ptx::umma_desc_sw128( sa + SL::A_BYTES );
#pragma unroll
						for( int k = 0; k < BK / 16; k++ )
						{
							// advance 16 f16 = 32 bytes along K inside the 128-byte swizzle row: +2 in the (addr >> 4) field
							ptx::umma_f16( tmem_d, da + (uint64_t)( k * 2 ), db + (uint64_t)( k * 2 ), idesc, ( kb | k ) != 0 ? 1u : 0u );
						}
						ptx::umma_commit( &bar_empty[ stage ] );
						if( kb == num_kb - 1 )
							ptx::umma_commit( &bar_tfull[ as ] );
						if( ++stage == STAGES ) { stage = 0; phase ^= 1; }
					}
					as ^= 1;
					if( as == 0 ) aphase ^= 1;
				}
			}
			__syncwarp();
		}
		else
		{
			// ===== epilogue warps =====
			const int q = warp & 3;   // TMEM lane quadrant accessible to this warp
			int as = 0;
			uint32_t aphase = 0;
			for( int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x )
			{
				const int m0 = ( tile / num_n ) * BM;
				const int n0 = ( tile % num_n ) * BN;
				// the residual of the first column chunk does not depend on the accumulator: request it before waiting for the MMAs
				float rsNext[ 32 ];
				if constexpr( MODE == EPI_BIAS_RESID ) loadResid( ep, m0 + q * 32, n0 + lane, rsNext );
				ptx::mbar_wait( &bar_tfull[ as ], aphase );
				ptx::tc_fence_after();
				const int m = m0 + q * 32 + lane;
				const uint32_t taddr = tmem_base + ( (uint32_t)( q * 32 ) << 16 ) + (uint32_t)( as * BN );
				float* xpose = xposeTiles[ warp - 2 ];
#pragma unroll 1
				for( int c = 0; c < BN / 32; c++ )
				{
					uint32_t r[ 32 ];
					ptx::tmem_ld_32x32( taddr + (uint32_t)( c * 32 ), r );
					ptx::tmem_ld_wait();
					const int nc = n0 + c * 32;
					// f32 outputs go through the transpose (measured: out-projection 99 -> 79 us); f16 outputs keep the row-per-thread
					// layout with 16-byte stores — 2-byte scalar stores, even though contiguous across the warp, were measured 2x slower
					// (fc1 144 -> 269 us), and the V^T band of the QKV GEMM is stored transposed, for which row-per-thread is the coalesced layout
					constexpr bool rowLayout = ( MODE == EPI_QKV ) || ( MODE == EPI_CONV1 ) || ( MODE == EPI_BIAS_GELU_F16 ) || ( MODE == EPI_CROSSKV );
					if( rowLayout )
					{
						float v[ 32 ];
#pragma unroll
						for( int i = 0; i < 32; i++ )
							v[ i ] = __uint_as_float( r[ i ] );
						epilogue<MODE>( ep, m, nc, v );
					}
					else
					{
#pragma unroll
						for( int i = 0; i < 32; i++ )
							xpose[ lane * 33 + i ] = __uint_as_float( r[ i ] );
						__syncwarp();
						float w[ 32 ];
#pragma unroll
						for( int rr = 0; rr < 32; rr++ )
							w[ rr ] = xpose[ rr * 33 + lane ];
						__syncwarp();
						if constexpr( MODE == EPI_BIAS_RESID )
						{
							float rs[ 32 ];
#pragma unroll
							for( int rr = 0; rr < 32; rr++ ) rs[ rr ] = rsNext[ rr ];
							if( c + 1 < BN / 32 ) loadResid( ep, m0 + q * 32, nc + 32 + lane, rsNext );   // next chunk's residual while this one is stored
							epilogueT<MODE>( ep, m0 + q * 32, nc + lane, w, rs );
						}
						else
							epilogueT<MODE>( ep, m0 + q * 32, nc + lane, w );
					}
				}
				ptx::tc_fence_before();
				__syncwarp();
				if( lane == 0 )
					ptx::mbar_arrive( &bar_tempty[ as ] );
				as ^= 1;
				if( as == 0 ) aphase ^= 1;
			}
		}

		ptx::tc_fence_before();
		__syncthreads();
		if( warp == 1 )
		{
			ptx::tc_fence_after();
			ptx::tmem_dealloc( tmem_base, TMEM_COLS );
		}
	}

	// ---------------------------------------------------------------------------------------------------------------
	// host side
	typedef CUresult ( *PFN_encodeTiled )( CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
		const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill );

	static PFN_encodeTiled getEncodeTiled()
	{
		static PFN_encodeTiled fn = nullptr;
		static std::once_flag once;
		std::call_once( once, []() {
			void* p = nullptr;
			cudaDriverEntryPointQueryResult qres;
			if( cudaGetDriverEntryPoint( "cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres ) == cudaSuccess && qres == cudaDriverEntryPointSuccess )
				fn = (PFN_encodeTiled)p;
		} );
		return fn;
	}

	bool makeMap2D( CUtensorMap* map, const void* base, uint64_t inner, uint64_t rows, uint64_t rowStrideBytes, uint32_t boxRows )
	{
		PFN_encodeTiled enc = getEncodeTiled();
		if( !enc ) return false;
		cuuint64_t dims[ 2 ] = { inner, rows };
		cuuint64_t strides[ 1 ] = { rowStrideBytes };
		cuuint32_t box[ 2 ] = { (cuuint32_t)BK, boxRows };
		cuuint32_t estr[ 2 ] = { 1, 1 };
		CUresult r = enc( map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>( base ), dims, strides, box, estr,
			CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE );
		if( r != CUDA_SUCCESS )
		{
			fprintf( stderr, "whisper_b200: cuTensorMapEncodeTiled failed (%d): inner=%llu rows=%llu stride=%llu box=%u\n", (int)r,
				(unsigned long long)inner, (unsigned long long)rows, (unsigned long long)rowStrideBytes, boxRows );
			return false;
		}
		return true;
	}

	template<int BN, int STAGES, int MODE, int AMODE>
	static cudaError_t launchT( const Launch& L, int numSMs, cudaStream_t stream )
	{
		using SL = SmemLayout<BN, STAGES>;
		auto kfn = gemm_tc_kernel<BN, STAGES, MODE, AMODE>;
		static kern::PerDeviceMax attr;
		{
			cudaError_t e = attr.raise( SL::TOTAL, [ & ]( size_t n ) { return cudaFuncSetAttribute( kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)n ); } );
			if( e != cudaSuccess ) return e;
		}
		const int num_tiles = ( ( L.M + BM - 1 ) / BM ) * ( ( L.N + BN - 1 ) / BN );
		const int grid = num_tiles < numSMs ? num_tiles : numSMs;
		if( grid <= 0 ) return cudaSuccess;
		kfn<<<grid, 192, SL::TOTAL, stream>>>( L.mapA, L.mapA2, L.mapB, L.M, L.N, L.K, L.kTap, L.ep );
		return cudaGetLastError();
	}

	template<int MODE, int AMODE>
	static cudaError_t launchBN( const Launch& L, int bn, int numSMs, cudaStream_t stream )
	{
		if( bn == 256 ) return launchT<256, 4, MODE, AMODE>( L, numSMs, stream );
		return launchT<128, 6, MODE, AMODE>( L, numSMs, stream );
	}

	cudaError_t launch( const Launch& L, EpiMode epi, AMode amode, int bn, int numSMs, cudaStream_t stream )
	{
		if( amode == A_CONV_S1 && epi == EPI_CONV1 ) return launchBN<EPI_CONV1, A_CONV_S1>( L, bn, numSMs, stream );
		if( amode == A_CONV_S2 && epi == EPI_CONV2 ) return launchBN<EPI_CONV2, A_CONV_S2>( L, bn, numSMs, stream );
		if( amode != A_PLAIN ) return cudaErrorInvalidValue;
		switch( epi )
		{
		case EPI_F32: return launchBN<EPI_F32, A_PLAIN>( L, bn, numSMs, stream );
		case EPI_QKV: return launchBN<EPI_QKV, A_PLAIN>( L, bn, numSMs, stream );
		case EPI_BIAS_RESID: return launchBN<EPI_BIAS_RESID, A_PLAIN>( L, bn, numSMs, stream );
		case EPI_BIAS_GELU_F16: return launchBN<EPI_BIAS_GELU_F16, A_PLAIN>( L, bn, numSMs, stream );
		case EPI_CROSSKV: return launchBN<EPI_CROSSKV, A_PLAIN>( L, bn, numSMs, stream );
		default: return cudaErrorInvalidValue;
		}
	}
}
