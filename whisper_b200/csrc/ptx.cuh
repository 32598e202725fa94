// Thin inline-PTX layer for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld) and
// the UMMA shared-memory + instruction descriptors.  Everything here is hand-written for B200; there is no other target.
//
// Descriptor bit layouts were cross-checked against the CUTLASS headers vendored in this image
// (cute/arch/mma_sm100_desc.hpp: SmemDescriptor, InstrDescriptor) — the layouts are hardware facts, the code is ours.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ptx
{
	__device__ __forceinline__ uint32_t smem_u32( const void* p ) { return (uint32_t)__cvta_generic_to_shared( p ); }

	__device__ __forceinline__ bool elect_one()
	{
		uint32_t pred = 0;
		asm volatile(
			"{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
			: "=r"( pred ) );
		return pred != 0;
	}

	// ---------------------------------------------------------------------------------------------
	// mbarrier
	__device__ __forceinline__ void mbar_init( uint64_t* bar, uint32_t count )
	{
		asm volatile( "mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"( smem_u32( bar ) ), "r"( count ) : "memory" );
	}
	__device__ __forceinline__ void fence_barrier_init() { asm volatile( "fence.mbarrier_init.release.cluster;" ::: "memory" ); }
	__device__ __forceinline__ void fence_proxy_async() { asm volatile( "fence.proxy.async.shared::cta;" ::: "memory" ); }
	__device__ __forceinline__ void mbar_expect_tx( uint64_t* bar, uint32_t bytes )
	{
		asm volatile( "mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"( smem_u32( bar ) ), "r"( bytes ) : "memory" );
	}
	__device__ __forceinline__ void mbar_arrive( uint64_t* bar )
	{
		asm volatile( "mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"( smem_u32( bar ) ) : "memory" );
	}
	__device__ __forceinline__ bool mbar_try_wait( uint64_t* bar, uint32_t parity )
	{
		uint32_t ok;
		asm volatile(
			"{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
			: "=r"( ok )
			: "r"( smem_u32( bar ) ), "r"( parity )
			: "memory" );
		return ok != 0;
	}
	// Bounded wait: a protocol bug must surface as a trap (launch failure), never as a hung GPU.
	__device__ __forceinline__ void mbar_wait( uint64_t* bar, uint32_t parity )
	{
		uint32_t spins = 0;
		while( !mbar_try_wait( bar, parity ) )
		{
			if( ++spins > ( 1u << 22 ) )
				__trap();
		}
	}

	// ---------------------------------------------------------------------------------------------
	// TMA tile loads, completion signalled on an mbarrier (complete_tx::bytes)
	__device__ __forceinline__ void tma_load_2d( void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1 )
	{
		asm volatile(
			"cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
			::"r"( smem_u32( dst ) ), "l"( map ), "r"( smem_u32( bar ) ), "r"( c0 ), "r"( c1 )
			: "memory" );
	}
	// 1-D bulk copy global -> shared by the TMA engine (no tensor map): one instruction moves up to the whole tile, completion is
	// counted in bytes on the mbarrier.  Addresses and size must be multiples of 16.
	__device__ __forceinline__ void bulk_load_1d( void* dst, const void* src, uint32_t bytes, uint64_t* bar )
	{
		asm volatile( "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"( smem_u32( dst ) ), "l"( src ),
			"r"( bytes ), "r"( smem_u32( bar ) )
			: "memory" );
	}
	__device__ __forceinline__ void prefetch_tensormap( const CUtensorMap* map )
	{
		asm volatile( "prefetch.tensormap [%0];" ::"l"( map ) : "memory" );
	}

	// ---------------------------------------------------------------------------------------------
	// tcgen05: tensor memory management
	__device__ __forceinline__ void tmem_alloc( uint32_t* dst_smem, uint32_t ncols )
	{
		asm volatile( "tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"( smem_u32( dst_smem ) ), "r"( ncols ) : "memory" );
	}
	__device__ __forceinline__ void tmem_relinquish() { asm volatile( "tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory" ); }
	__device__ __forceinline__ void tmem_dealloc( uint32_t taddr, uint32_t ncols )
	{
		asm volatile( "tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"( taddr ), "r"( ncols ) : "memory" );
	}
	__device__ __forceinline__ void tc_fence_before() { asm volatile( "tcgen05.fence::before_thread_sync;" ::: "memory" ); }
	__device__ __forceinline__ void tc_fence_after() { asm volatile( "tcgen05.fence::after_thread_sync;" ::: "memory" ); }

	// D[tmem] (+)= A[smem desc] * B[smem desc], f16 inputs, f32 accumulate.  One thread issues for the CTA.
	__device__ __forceinline__ void umma_f16( uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate )
	{
		asm volatile(
			"{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
			"tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
			::"r"( tmem_d ), "l"( desc_a ), "l"( desc_b ), "r"( idesc ), "r"( accumulate )
			: "memory" );
	}
	// Arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed.
	// (implies tcgen05.fence::before_thread_sync)
	__device__ __forceinline__ void umma_commit( uint64_t* bar )
	{
		asm volatile( "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"( smem_u32( bar ) ) : "memory" );
	}

	// TMEM -> registers: 32 lanes (this warp's quadrant) x 32 consecutive 32-bit columns; thread i gets lane i.
	__device__ __forceinline__ void tmem_ld_32x32( uint32_t taddr, uint32_t* r )
	{
		asm volatile(
			"tcgen05.ld.sync.aligned.32x32b.x32.b32 "
			"{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
			"%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
			: "=r"( r[ 0 ] ), "=r"( r[ 1 ] ), "=r"( r[ 2 ] ), "=r"( r[ 3 ] ), "=r"( r[ 4 ] ), "=r"( r[ 5 ] ), "=r"( r[ 6 ] ), "=r"( r[ 7 ] ),
			  "=r"( r[ 8 ] ), "=r"( r[ 9 ] ), "=r"( r[ 10 ] ), "=r"( r[ 11 ] ), "=r"( r[ 12 ] ), "=r"( r[ 13 ] ), "=r"( r[ 14 ] ), "=r"( r[ 15 ] ),
			  "=r"( r[ 16 ] ), "=r"( r[ 17 ] ), "=r"( r[ 18 ] ), "=r"( r[ 19 ] ), "=r"( r[ 20 ] ), "=r"( r[ 21 ] ), "=r"( r[ 22 ] ), "=r"( r[ 23 ] ),
			  "=r"( r[ 24 ] ), "=r"( r[ 25 ] ), "=r"( r[ 26 ] ), "=r"( r[ 27 ] ), "=r"( r[ 28 ] ), "=r"( r[ 29 ] ), "=r"( r[ 30 ] ), "=r"( r[ 31 ] )
			: "r"( taddr )
			: "memory" );
	}
	__device__ __forceinline__ void tmem_ld_wait() { asm volatile( "tcgen05.wait::ld.sync.aligned;" ::: "memory" ); }
	// registers -> TMEM, the mirror image of tmem_ld_32x32
	__device__ __forceinline__ void tmem_st_32x32( uint32_t taddr, const uint32_t* r )
	{
		asm volatile(
			"tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
			"{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
			"%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
			:
			: "r"( taddr ), "r"( r[ 0 ] ), "r"( r[ 1 ] ), "r"( r[ 2 ] ), "r"( r[ 3 ] ), "r"( r[ 4 ] ), "r"( r[ 5 ] ), "r"( r[ 6 ] ), "r"( r[ 7 ] ),
			  "r"( r[ 8 ] ), "r"( r[ 9 ] ), "r"( r[ 10 ] ), "r"( r[ 11 ] ), "r"( r[ 12 ] ), "r"( r[ 13 ] ), "r"( r[ 14 ] ), "r"( r[ 15 ] ),
			  "r"( r[ 16 ] ), "r"( r[ 17 ] ), "r"( r[ 18 ] ), "r"( r[ 19 ] ), "r"( r[ 20 ] ), "r"( r[ 21 ] ), "r"( r[ 22 ] ), "r"( r[ 23 ] ),
			  "r"( r[ 24 ] ), "r"( r[ 25 ] ), "r"( r[ 26 ] ), "r"( r[ 27 ] ), "r"( r[ 28 ] ), "r"( r[ 29 ] ), "r"( r[ 30 ] ), "r"( r[ 31 ] )
			: "memory" );
	}
	__device__ __forceinline__ void tmem_st_wait() { asm volatile( "tcgen05.wait::st.sync.aligned;" ::: "memory" ); }

	// ---------------------------------------------------------------------------------------------
	// UMMA descriptors
	//
	// Shared-memory matrix descriptor for a K-major operand tile stored as rows of 64 f16 (128 bytes) with the
	// 128-byte swizzle TMA writes (CU_TENSOR_MAP_SWIZZLE_128B): 8-row groups are 1024 bytes apart (SBO), LBO is
	// unused for swizzled K-major (encoded 1), version = 1 (sm_100), layout type 2 = SWIZZLE_128B.
	__device__ __forceinline__ uint64_t umma_desc_sw128( uint32_t smem_addr )
	{
		uint64_t d = 0;
		d |= (uint64_t)( ( smem_addr & 0x3FFFF ) >> 4 );         // [0,14)  start address >> 4
		d |= (uint64_t)1 << 16;                                   // [16,30) leading byte offset >> 4 (ignored)
		d |= (uint64_t)( 1024 >> 4 ) << 32;                       // [32,46) stride byte offset >> 4
		d |= (uint64_t)1 << 46;                                   // [46,48) descriptor version (Blackwell)
		d |= (uint64_t)2 << 61;                                   // [61,64) SWIZZLE_128B
		return d;
	}
	// Instruction descriptor, kind::f16: A,B = f16 (format 0), D = f32 (1), both operands K-major, dense.
	__host__ __device__ constexpr uint32_t umma_idesc_f16( int M, int N )
	{
		return ( 1u << 4 ) | ( (uint32_t)( N >> 3 ) << 17 ) | ( (uint32_t)( M >> 4 ) << 24 );
	}

	// Programmatic dependent launch: let the next kernel of the stream start its prologue (weight / KV prefetch) while this one runs,
	// and block until every kernel this one depends on has completed and flushed its writes.
	__device__ __forceinline__ void pdl_launch_dependents() { asm volatile( "griddepcontrol.launch_dependents;" ::: "memory" ); }
	__device__ __forceinline__ void pdl_wait() { asm volatile( "griddepcontrol.wait;" ::: "memory" ); }

	__device__ __forceinline__ float gelu_f16_semantics( float x )
	{
		// Oracle: y = f32( T[f16(x)] ), T[i] = f16( gelu( f32(i) ) ) evaluated in double (ggml.c:999-1021, table init :1372-1383).
		// Both roundings are reproduced; the table is replaced by evaluating the same function as v * sigmoid(2u),
		// u = sqrt(2/pi) * v * (1 + 0.044715 v^2)  ==  0.5 v (1 + tanh(u)), with ex2/rcp (relative error ~1e-6, i.e. the f16 result
		// differs from the table only on near-ties).  This form costs ~10 instructions instead of tanhf's ~30: the fc1 epilogue
		// evaluates it 49 M times per encoder layer.
		const float v = __half2float( __float2half_rn( x ) );
		const float u2 = -1.5957691216057308f * v * ( 1.0f + 0.044715f * v * v );   // -2u
		const float y = __fdividef( v, 1.0f + __expf( u2 ) );
		return __half2float( __float2half_rn( y ) );
	}
}
