// Context: workspaces, tensor maps, encoder / decoder orchestration (the network definition).
// Reference: WhisperContext::encode / ::decode (Whisper/Whisper/WhisperContext.cpp:310-388, 578-639); oracle whisper_encode /
// whisper_decode (Whisper/source/whisper.cpp:1084-1872).  All launches go to one stream; the N = 1 decoder step is a CUDA graph.
#include "engine.h"
#include <math.h>
#include <memory>
#include <stdlib.h>

namespace wsp
{
	namespace
	{
		thread_local uint64_t t_allocated = 0;   // bytes handed out by devAlloc on this thread (read by its callers)
		template<class T> int devAlloc( T*& p, size_t count, bool zero = false )
		{
			void* v = nullptr;
			cudaError_t e = cudaMalloc( &v, count * sizeof( T ) );
			if( e != cudaSuccess ) return cudaFail( e, "cudaMalloc" );
			if( zero )
			{
				e = cudaMemset( v, 0, count * sizeof( T ) );
				if( e != cudaSuccess ) return cudaFail( e, "cudaMemset" );
			}
			p = static_cast<T*>( v );
			t_allocated += count * sizeof( T );
			return WSP_OK;
		}
		inline void launched( int n = 1 ) { g_launchCount.fetch_add( (uint64_t)n, std::memory_order_relaxed ); }
		inline int pickBN( int N ) { return ( N % 256 == 0 ) ? 256 : 128; }
	}

	Context::~Context()
	{
		if( e ) cudaSetDevice( e->device );
		if( stepGraph ) cudaGraphExecDestroy( stepGraph );
		for( auto& s : slots ) { if( s.mel ) cudaFree( s.mel ); if( s.pcm ) cudaFree( s.pcm ); }
		for( auto& v : prof.pool ) cudaEventDestroy( v );
		for( auto& v : timerEv ) if( v ) cudaEventDestroy( v );
		void* bufs[] = { flowLayers, flowBias, flowExch, flowCtrl, megaLayers, megaBarrier, megaTiming, melMax, pcmDev, melF16, conv1, x, xn, q, k, vt, attn, h, crossK, crossV, selfK, selfV, xd, qd, attnD, hD, logits, probs, tieScratch,
			tokensDev, dNPast, sampled, history };
		for( void* b : bufs ) if( b ) cudaFree( b );
		for( auto& v : ev ) if( v ) cudaEventDestroy( v );
		if( stream ) cudaStreamDestroy( stream );
	}

	int createContext( Engine* e, int maxBatch, Context** out )
	{
		if( !e || maxBatch < 1 || maxBatch > 256 ) return fail( WSP_E_INVALIDARG, "max_batch must be in [1, 256]" );
		WSP_CUDA( cudaSetDevice( e->device ) );
		std::unique_ptr<Context> cp( new Context() );
		Context& c = *cp;
		c.e = e;
		c.maxB = maxBatch;
		t_allocated = 0;
		struct Tally { Context& c; ~Tally() { c.devBytes += t_allocated; t_allocated = 0; } } tally{ c };
		const HParams& hp = e->hp;
		const int d = hp.n_audio_state, H = hp.n_audio_head, T = hp.n_audio_ctx, L = hp.n_text_layer;
		if( T * 2 != kFrames ) return fail( WSP_E_FORMAT, "n_audio_ctx must be 1500" );
		c.Tp = ( ( T + 127 ) / 128 ) * 128;
		WSP_CUDA( cudaStreamCreateWithFlags( &c.stream, cudaStreamNonBlocking ) );
		for( auto& v : c.ev ) WSP_CUDA( cudaEventCreate( &v ) );
		for( auto& v : c.timerEv ) WSP_CUDA( cudaEventCreate( &v ) );
		c.slots.resize( maxBatch );
		const size_t B = (size_t)maxBatch;
		WSP_CHECK( devAlloc( c.melMax, B ) );
		WSP_CHECK( devAlloc( c.melF16, B * kFramesPad * 80, true ) );          // halo rows stay zero forever
		WSP_CHECK( devAlloc( c.conv1, B * kFramesPad * d, true ) );
		WSP_CHECK( devAlloc( c.x, B * T * d ) );
		WSP_CHECK( devAlloc( c.xn, B * T * d ) );
		WSP_CHECK( devAlloc( c.q, B * T * d ) );
		WSP_CHECK( devAlloc( c.k, B * T * d ) );
		WSP_CHECK( devAlloc( c.vt, B * H * 64 * c.Tp, true ) );                // pad columns [T, Tp) stay zero
		WSP_CHECK( devAlloc( c.attn, B * T * d ) );
		WSP_CHECK( devAlloc( c.h, B * T * 4 * d ) );
		WSP_CHECK( devAlloc( c.crossK, (size_t)L * B * T * d ) );
		WSP_CHECK( devAlloc( c.crossV, (size_t)L * B * T * d ) );
		WSP_CHECK( devAlloc( c.selfK, (size_t)L * B * hp.n_text_ctx * d, true ) );
		WSP_CHECK( devAlloc( c.selfV, (size_t)L * B * hp.n_text_ctx * d, true ) );
		WSP_CHECK( devAlloc( c.xd, B * kMaxDecodeTokens * d ) );
		WSP_CHECK( devAlloc( c.qd, B * kMaxDecodeTokens * d ) );
		WSP_CHECK( devAlloc( c.attnD, B * kMaxDecodeTokens * d ) );
		WSP_CHECK( devAlloc( c.hD, B * kMaxDecodeTokens * 4 * d ) );
		WSP_CHECK( devAlloc( c.logits, B * kAllLogitsTokens * hp.n_vocab ) );
		WSP_CHECK( devAlloc( c.probs, B * kAllLogitsTokens * hp.n_vocab ) );
		WSP_CHECK( devAlloc( c.tieScratch, B * ( (size_t)hp.n_vocab + 1024 ) ) );
		WSP_CHECK( devAlloc( c.tokensDev, B * kMaxDecodeTokens, true ) );
		WSP_CHECK( devAlloc( c.dNPast, 8, true ) );
		c.dFlags = c.dNPast + 2;
		c.dStep = c.dNPast + 4;
		WSP_CHECK( devAlloc( c.sampled, B ) );
		c.histCap = hp.n_text_ctx;
		WSP_CHECK( devAlloc( c.history, B * c.histCap, true ) );

		// ---- tensor maps ----
		c.bnD = pickBN( d ); c.bn3D = pickBN( 3 * d ); c.bn4D = pickBN( 4 * d ); c.bnCross = pickBN( L * 2 * d );
		bool ok = true;
		// A operands (activations): box rows = 128
		ok &= gemm::makeMap2D( &c.mapMel, c.melF16, 80, B * kFramesPad, 80 * 2, gemm::BM );
		ok &= gemm::makeMap2D( &c.mapConv1Even, c.conv1, d, B * ( kFramesPad / 2 ), (uint64_t)2 * d * 2, gemm::BM );
		ok &= gemm::makeMap2D( &c.mapConv1Odd, c.conv1 + d, d, B * ( kFramesPad / 2 ), (uint64_t)2 * d * 2, gemm::BM );
		ok &= gemm::makeMap2D( &c.mapXn, c.xn, d, B * T, (uint64_t)d * 2, gemm::BM );
		ok &= gemm::makeMap2D( &c.mapAttn, c.attn, d, B * T, (uint64_t)d * 2, gemm::BM );
		ok &= gemm::makeMap2D( &c.mapH, c.h, 4 * d, B * T, (uint64_t)4 * d * 2, gemm::BM );
		// B operands (weights): box rows = BN of the GEMM that uses them
		ok &= gemm::makeMap2D( &c.mapConv1W, e->conv1w, 3 * kConv1KTap, d, (uint64_t)3 * kConv1KTap * 2, c.bnD );
		ok &= gemm::makeMap2D( &c.mapConv2W, e->conv2w, 3 * d, d, (uint64_t)3 * d * 2, c.bnD );
		ok &= gemm::makeMap2D( &c.mapCrossW, e->crossW, d, (uint64_t)L * 2 * d, (uint64_t)d * 2, c.bnCross );
		const int Le = hp.n_audio_layer;
		c.mapWqkv.resize( Le ); c.mapWo.resize( Le ); c.mapW1.resize( Le ); c.mapW2.resize( Le );
		for( int i = 0; i < Le; i++ )
		{
			ok &= gemm::makeMap2D( &c.mapWqkv[ i ], e->enc[ i ].wqkv, d, 3 * d, (uint64_t)d * 2, c.bn3D );
			ok &= gemm::makeMap2D( &c.mapWo[ i ], e->enc[ i ].wo, d, d, (uint64_t)d * 2, c.bnD );
			ok &= gemm::makeMap2D( &c.mapW1[ i ], e->enc[ i ].w1, d, 4 * d, (uint64_t)d * 2, c.bn4D );
			ok &= gemm::makeMap2D( &c.mapW2[ i ], e->enc[ i ].w2, 4 * d, d, (uint64_t)4 * d * 2, c.bnD );
		}
		// attention operands
		ok &= gemm::makeMap2D( &c.mapQ, c.q, 64, B * H * T, 128, 128 );
		ok &= gemm::makeMap2D( &c.mapK, c.k, 64, B * H * T, 128, 128 );
		ok &= gemm::makeMap2D( &c.mapVt, c.vt, c.Tp, B * H * 64, (uint64_t)c.Tp * 2, 64 );
		if( !ok ) return fail( WSP_E_CUDA, "cuTensorMapEncodeTiled failed (driver too old for TMA?)" );
		// per-layer pointer table for the persistent decoder-step kernel
		{
			std::vector<kern::MegaLayer> ml( (size_t)L );
			for( int i = 0; i < L; i++ )
			{
				const DecLayerW& W = e->dec[ i ];
				kern::MegaLayer& m = ml[ i ];
				m.ln1g = W.ln1.g; m.ln1b = W.ln1.b; m.lncg = W.lnc.g; m.lncb = W.lnc.b; m.ln3g = W.ln3.g; m.ln3b = W.ln3.b;
				m.wqkv = W.wqkv; m.wo = W.wo; m.wcq = W.wcq; m.wco = W.wco; m.w1 = W.w1; m.w2 = W.w2;
				m.bqkv = W.bqkv; m.bo = W.bo; m.bcq = W.bcq; m.bco = W.bco; m.b1 = W.b1; m.b2 = W.b2;
				m.kCache = c.selfK + (size_t)i * B * hp.n_text_ctx * d;
				m.vCache = c.selfV + (size_t)i * B * hp.n_text_ctx * d;
				m.crossK = c.crossK + (size_t)i * B * T * d;
				m.crossV = c.crossV + (size_t)i * B * T * d;
			}
			WSP_CHECK( devAlloc( c.megaLayers, (size_t)L ) );
			WSP_CUDA( cudaMemcpy( c.megaLayers, ml.data(), ml.size() * sizeof( kern::MegaLayer ), cudaMemcpyHostToDevice ) );
			WSP_CHECK( devAlloc( c.megaBarrier, 64, true ) );
			WSP_CHECK( devAlloc( c.megaTiming, 4608, true ) );
		}
		// dataflow decoder-step kernel: pointer table, per-CTA bias slabs, sentinel-filled exchange buffers (decode_flow.cu)
		{
			const kern::FlowGeom g = kern::flowGeometry( d, hp.n_vocab, e->numSMs );
			c.flowGeom = g;
			const size_t slabLayer = (size_t)g.grid * g.slabFloats;
			WSP_CHECK( devAlloc( c.flowBias, slabLayer * L ) );
			std::vector<kern::FlowLayer> fl( (size_t)L );
			for( int i = 0; i < L; i++ )
			{
				const DecLayerW& W = e->dec[ i ];
				kern::FlowLayer& m = fl[ i ];
				m.ln1g = W.ln1.g; m.ln1b = W.ln1.b; m.lncg = W.lnc.g; m.lncb = W.lnc.b; m.ln3g = W.ln3.g; m.ln3b = W.ln3.b;
				m.wqkv = W.wqkv; m.wo = W.wo; m.wcq = W.wcq; m.wco = W.wco; m.w1 = W.w1; m.w2 = W.w2;
				m.biasSlab = c.flowBias + (size_t)i * slabLayer;
				m.kCache = c.selfK + (size_t)i * B * hp.n_text_ctx * d;
				m.vCache = c.selfV + (size_t)i * B * hp.n_text_ctx * d;
				m.crossK = c.crossK + (size_t)i * B * T * d;
				m.crossV = c.crossV + (size_t)i * B * T * d;
				WSP_CUDA( kern::flowBuildBiasSlab( c.flowBias + (size_t)i * slabLayer, g, d, W.bqkv, W.bo, W.bcq, W.bco, W.b1, W.b2, c.stream ) );
			}
			WSP_CHECK( devAlloc( c.flowLayers, (size_t)L ) );
			WSP_CUDA( cudaMemcpy( c.flowLayers, fl.data(), fl.size() * sizeof( kern::FlowLayer ), cudaMemcpyHostToDevice ) );
			const size_t exBytes = kern::flowExchangeBytes( d, maxBatch, L );
			WSP_CHECK( devAlloc( c.flowExch, exBytes ) );
			WSP_CUDA( cudaMemset( c.flowExch, 0xFF, exBytes ) );
			WSP_CHECK( devAlloc( c.flowCtrl, 8, true ) );
			const char* env = getenv( "WSP_STEP_MODE" );
			if( env && env[ 0 ] >= '0' && env[ 0 ] <= '2' ) c.stepMode = env[ 0 ] - '0';
			env = getenv( "WSP_TIMING_CTA" );
			if( env ) c.stepTimingCta = atoi( env );
		}
		WSP_CUDA( kern::prepare( 4 * d ) );
		WSP_CUDA( kern::megaPrepare( d ) );
		WSP_CUDA( kern::flowPrepare( d ) );
		WSP_CUDA( cudaDeviceSynchronize() );
		*out = cp.release();
		return WSP_OK;
	}

	// ===============================================================================================================
	// mel
	// ===============================================================================================================
	static int ensureMelSlot( Context& c, int slot, int nLen )
	{
		if( slot < 0 || slot >= c.maxB ) return fail( WSP_E_BOUNDS, "chunk slot out of range" );
		MelSlot& s = c.slots[ slot ];
		if( nLen > s.cap )
		{
			if( s.mel ) cudaFree( s.mel );
			s.mel = nullptr;
			const int cap = nLen < kFrames ? kFrames : nLen;
			c.devBytes -= (uint64_t)80 * s.cap * 4;
			WSP_CHECK( devAlloc( s.mel, (size_t)80 * cap ) );
			c.devBytes += (uint64_t)80 * cap * 4;
			s.cap = cap;
		}
		s.nLen = nLen;
		return WSP_OK;
	}

	static int melFromDevicePcm( Context& c, int slot, const float* pcmDev, int nSamples )
	{
		const int nLen = nSamples / 160;   // whisper.cpp:2080
		WSP_CHECK( ensureMelSlot( c, slot, nLen ) );
		MelSlot& s = c.slots[ slot ];
		WSP_CUDA( kern::melPower( c.e->mel, pcmDev, nSamples, nLen, s.mel, c.melMax + slot, c.stream ) );
		WSP_CUDA( kern::melNormalize( s.mel, 80 * nLen, c.melMax + slot, c.stream ) );
		launched( 3 );
		return WSP_OK;
	}

	// log-mel of chunks [0, batch) in one pair of launches per 16 chunks: pcmOf( b ) / samplesOf( b ) name chunk b's device PCM
	template<class P, class N>
	static int melBatchFromDevicePcm( Context& c, int batch, P pcmOf, N samplesOf )
	{
		for( int b0 = 0; b0 < batch; b0 += kern::MEL_BATCH )
		{
			kern::MelBatch mb{};
			mb.count = batch - b0 < kern::MEL_BATCH ? batch - b0 : kern::MEL_BATCH;
			mb.maxSlots = c.melMax + b0;
			for( int i = 0; i < mb.count; i++ )
			{
				const int n = samplesOf( b0 + i );
				WSP_CHECK( ensureMelSlot( c, b0 + i, n / 160 ) );   // whisper.cpp:2080
				mb.pcm[ i ] = pcmOf( b0 + i );
				mb.mel[ i ] = c.slots[ b0 + i ].mel;
				mb.nSamples[ i ] = n;
				mb.nLen[ i ] = n / 160;
			}
			WSP_CUDA( kern::melBatch( c.e->mel, mb, c.stream ) );
			launched( 3 );
		}
		return WSP_OK;
	}

	static int ensurePcm( Context& c, size_t floats )
	{
		if( floats <= c.pcmCap ) return WSP_OK;
		if( c.pcmDev ) { WSP_CUDA( cudaStreamSynchronize( c.stream ) ); cudaFree( c.pcmDev ); c.pcmDev = nullptr; }
		c.devBytes -= (uint64_t)c.pcmCap * 4;
		WSP_CHECK( devAlloc( c.pcmDev, floats ) );
		c.devBytes += (uint64_t)floats * 4;
		c.pcmCap = floats;
		return WSP_OK;
	}

	int ctxPcmToMel( Context& c, int slot, const float* pcmHost, int nSamples )
	{
		if( ( !pcmHost && nSamples != 0 ) || nSamples < 0 ) return fail( WSP_E_INVALIDARG, "pcm" );
		WSP_CUDA( cudaSetDevice( c.e->device ) );
		WSP_CHECK( ensurePcm( c, (size_t)( nSamples > 0 ? nSamples : 1 ) ) );
		WSP_CUDA( cudaEventRecord( c.ev[ 0 ], c.stream ) );
		if( nSamples > 0 )
			WSP_CUDA( cudaMemcpyAsync( c.pcmDev, pcmHost, (size_t)nSamples * 4, cudaMemcpyHostToDevice, c.stream ) );
		WSP_CHECK( melFromDevicePcm( c, slot, c.pcmDev, nSamples ) );
		WSP_CUDA( cudaEventRecord( c.ev[ 1 ], c.stream ) );
		WSP_CUDA( cudaStreamSynchronize( c.stream ) );
		float ms = 0;
		cudaEventElapsedTime( &ms, c.ev[ 0 ], c.ev[ 1 ] );
		c.ms[ 0 ] += ms; c.calls[ 0 ]++;
		return WSP_OK;
	}

	// One window of a streamed clip (Whisper/Whisper/MelStreamer.cpp:128-236): frames [0, nFrames) of the PCM handed in — the PCM may
	// (and, except at the end of a stream, does) extend past the last frame so that every frame sees its full 400 samples — normalised
	// by the maximum over THESE frames only, floored at 1e-20 (:136), or by `forcedMax` when the streamer re-uses the maximum of the
	// previous, longer window that ended at the same frame (:152-166).  Same kernels as the whole-clip path.
	int ctxPcmToMelWindow( Context& c, int slot, const float* pcmHost, int nSamples, int nFrames, const float* forcedMax, float* maxOut )
	{
		if( ( !pcmHost && nSamples != 0 ) || nSamples < 0 || nFrames < 0 ) return fail( WSP_E_INVALIDARG, "pcm / frame count" );
		WSP_CUDA( cudaSetDevice( c.e->device ) );
		WSP_CHECK( ensurePcm( c, (size_t)( nSamples > 0 ? nSamples : 1 ) ) );
		WSP_CHECK( ensureMelSlot( c, slot, nFrames ) );
		MelSlot& s = c.slots[ slot ];
		WSP_CUDA( cudaEventRecord( c.ev[ 0 ], c.stream ) );
		if( nSamples > 0 )
			WSP_CUDA( cudaMemcpyAsync( c.pcmDev, pcmHost, (size_t)nSamples * 4, cudaMemcpyHostToDevice, c.stream ) );
		WSP_CUDA( kern::melPower( c.e->mel, c.pcmDev, nSamples, nFrames, s.mel, c.melMax + slot, c.stream ) );
		int32_t ordered = 0;
		WSP_CUDA( cudaMemcpyAsync( &ordered, c.melMax + slot, 4, cudaMemcpyDeviceToHost, c.stream ) );
		WSP_CUDA( cudaStreamSynchronize( c.stream ) );
		// the device keeps the maximum as an order-preserving integer (kernels_frontend.cu: orderedFromFloat)
		auto toFloat = []( int32_t i ) { const int32_t b = i >= 0 ? i : i ^ 0x7FFFFFFF; float f; memcpy( &f, &b, 4 ); return f; };
		auto toOrdered = []( float f ) { int32_t b; memcpy( &b, &f, 4 ); return b >= 0 ? b : b ^ 0x7FFFFFFF; };
		float mmax = toFloat( ordered );
		if( !( mmax > 1e-20f ) ) mmax = 1e-20f;
		if( maxOut ) *maxOut = mmax;
		if( forcedMax ) mmax = *forcedMax;
		const int32_t wanted = toOrdered( mmax );
		if( wanted != ordered )
			WSP_CUDA( cudaMemcpyAsync( c.melMax + slot, &wanted, 4, cudaMemcpyHostToDevice, c.stream ) );
		WSP_CUDA( kern::melNormalize( s.mel, 80 * nFrames, c.melMax + slot, c.stream ) );
		launched( 3 );
		WSP_CUDA( cudaEventRecord( c.ev[ 1 ], c.stream ) );
		WSP_CUDA( cudaStreamSynchronize( c.stream ) );
		float ms = 0;
		cudaEventElapsedTime( &ms, c.ev[ 0 ], c.ev[ 1 ] );
		c.ms[ 0 ] += ms; c.calls[ 0 ]++;
		return WSP_OK;
	}

	int ctxSetMel( Context& c, int slot, const float* melHost, int nLen )
	{
		if( !melHost || nLen < 0 ) return fail( WSP_E_INVALIDARG, "mel" );
		WSP_CUDA( cudaSetDevice( c.e->device ) );
		WSP_CHECK( ensureMelSlot( c, slot, nLen ) );
		WSP_CUDA( cudaMemcpyAsync( c.slots[ slot ].mel, melHost, (size_t)80 * nLen * 4, cudaMemcpyHostToDevice, c.stream ) );
		WSP_CUDA( cudaStreamSynchronize( c.stream ) );
		return WSP_OK;
	}

	// ===============================================================================================================
	// encoder
	// ===============================================================================================================
	static int encodeAsync( Context& c, const int32_t* offsets, int batch )
	{
		Engine& e = *c.e;
		const HParams& hp = e.hp;
		const int d = hp.n_audio_state, H = hp.n_audio_head, T = hp.n_audio_ctx, L = hp.n_text_layer;
		cudaStream_t s = c.stream;
		for( int b = 0; b < batch; b++ )
		{
			const MelSlot& ms = c.slots[ b ];
			if( !ms.mel ) return fail( WSP_E_INVALIDARG, "no mel in slot " + std::to_string( b ) );
			const int off = offsets ? offsets[ b ] : 0;
			if( off < 0 ) return fail( WSP_E_INVALIDARG, "negative mel offset" );
			WSP_CUDA( kern::melWindow( ms.mel, ms.nLen, off, c.melF16 + (size_t)b * kFramesPad * 80, kFrames, s ) );
			launched();
		}
		gemm::Launch g;
		// conv1 + bias + GELU -> f16, time-major with halo (a3)
		g = gemm::Launch();
		g.mapA = c.mapMel; g.mapA2 = c.mapMel; g.mapB = c.mapConv1W;
		g.M = batch * kFramesPad - 2; g.N = d; g.K = 3 * kConv1KTap; g.kTap = kConv1KTap / gemm::BK;
		g.ep.M = g.M; g.ep.N = d; g.ep.ld = d; g.ep.bias = e.conv1b; g.ep.out_a = c.conv1;
		g.ep.rows_per_chunk = kFramesPad; g.ep.valid_per_chunk = kFrames; g.ep.nchunks = batch;
		WSP_CUDA( gemm::launch( g, gemm::EPI_CONV1, gemm::A_CONV_S1, c.bnD, e.numSMs, s ) );
		// conv2 (stride 2) + bias + GELU + positional embedding -> f32 residual stream (a4)
		g = gemm::Launch();
		g.mapA = c.mapConv1Even; g.mapA2 = c.mapConv1Odd; g.mapB = c.mapConv2W;
		g.M = batch * ( kFramesPad / 2 ) - 1; g.N = d; g.K = 3 * d; g.kTap = d / gemm::BK;
		g.ep.M = g.M; g.ep.N = d; g.ep.ld = d; g.ep.bias = e.conv2b; g.ep.pos = e.encPos; g.ep.out_f32 = c.x;
		g.ep.rows_per_chunk = kFramesPad / 2; g.ep.valid_per_chunk = T; g.ep.nchunks = batch; g.ep.T = T; g.ep.d = d;
		WSP_CUDA( gemm::launch( g, gemm::EPI_CONV2, gemm::A_CONV_S2, c.bnD, e.numSMs, s ) );
		launched( 2 );

		const int M = batch * T;
		const int nLayers = ( c.debugEncLayers >= 0 && c.debugEncLayers < hp.n_audio_layer ) ? c.debugEncLayers : hp.n_audio_layer;
		for( int il = 0; il < nLayers; il++ )
		{
			const EncLayerW& W = e.enc[ il ];
			WSP_CUDA( kern::layerNormF16( c.x, W.ln1.g, W.ln1.b, c.xn, M, d, s ) );
			// Q | K | V projections in one GEMM; epilogue writes head-major f16 Q, K and transposed V (a6, a7)
			g = gemm::Launch();
			g.mapA = c.mapXn; g.mapA2 = c.mapXn; g.mapB = c.mapWqkv[ il ];
			g.M = M; g.N = 3 * d; g.K = d;
			g.ep.M = M; g.ep.N = 3 * d; g.ep.bias = W.bqkv; g.ep.out_a = c.q; g.ep.out_b = c.k; g.ep.out_c = c.vt;
			g.ep.T = T; g.ep.Tp = c.Tp; g.ep.H = H; g.ep.d = d;
			WSP_CUDA( gemm::launch( g, gemm::EPI_QKV, gemm::A_PLAIN, c.bn3D, e.numSMs, s ) );
			// fused flash attention (a8)
			attn::EncParams ap;
			ap.T = T; ap.H = H; ap.nBH = batch * H; ap.d = d; ap.out = c.attn;
			ap.scale_log2 = (float)( ( 1.0 / sqrt( 64.0 ) ) * 1.4426950408889634 );
			WSP_CUDA( attn::launchEnc( c.mapQ, c.mapK, c.mapVt, ap, s ) );
			// output projection + bias + residual (a9)
			g = gemm::Launch();
			g.mapA = c.mapAttn; g.mapA2 = c.mapAttn; g.mapB = c.mapWo[ il ];
			g.M = M; g.N = d; g.K = d;
			g.ep.M = M; g.ep.N = d; g.ep.ld = d; g.ep.bias = W.bo; g.ep.resid = c.x; g.ep.out_f32 = c.x;
			WSP_CUDA( gemm::launch( g, gemm::EPI_BIAS_RESID, gemm::A_PLAIN, c.bnD, e.numSMs, s ) );
			// MLP (a10)
			WSP_CUDA( kern::layerNormF16( c.x, W.ln2.g, W.ln2.b, c.xn, M, d, s ) );
			g = gemm::Launch();
			g.mapA = c.mapXn; g.mapA2 = c.mapXn; g.mapB = c.mapW1[ il ];
			g.M = M; g.N = 4 * d; g.K = d;
			g.ep.M = M; g.ep.N = 4 * d; g.ep.ld = 4 * d; g.ep.bias = W.b1; g.ep.out_a = c.h;
			WSP_CUDA( gemm::launch( g, gemm::EPI_BIAS_GELU_F16, gemm::A_PLAIN, c.bn4D, e.numSMs, s ) );
			g = gemm::Launch();
			g.mapA = c.mapH; g.mapA2 = c.mapH; g.mapB = c.mapW2[ il ];
			g.M = M; g.N = d; g.K = 4 * d;
			g.ep.M = M; g.ep.N = d; g.ep.ld = d; g.ep.bias = W.b2; g.ep.resid = c.x; g.ep.out_f32 = c.x;
			WSP_CUDA( gemm::launch( g, gemm::EPI_BIAS_RESID, gemm::A_PLAIN, c.bnD, e.numSMs, s ) );
			launched( 7 );
		}
		// ln_post (a11) and the cross-attention K/V memories of every decoder layer in one GEMM (a12)
		WSP_CUDA( kern::layerNormF16( c.x, e.encLnPost.g, e.encLnPost.b, c.xn, M, d, s ) );
		g = gemm::Launch();
		g.mapA = c.mapXn; g.mapA2 = c.mapXn; g.mapB = c.mapCrossW;
		g.M = M; g.N = L * 2 * d; g.K = d;
		g.ep.M = M; g.ep.N = g.N; g.ep.bias = e.crossB; g.ep.out_a = c.crossK; g.ep.out_b = c.crossV;
		g.ep.T = T; g.ep.H = H; g.ep.d = d; g.ep.nchunks = c.maxB;
		g.ep.scale = (float)pow( 64.0, -0.25 );   // whisper.cpp:1465
		WSP_CUDA( gemm::launch( g, gemm::EPI_CROSSKV, gemm::A_PLAIN, c.bnCross, e.numSMs, s ) );
		launched( 2 );
		return WSP_OK;
	}

	int ctxEncode( Context& c, const int32_t* offsets, int batch )
	{
		if( batch < 1 || batch > c.maxB ) return fail( WSP_E_BOUNDS, "batch exceeds the context's max_batch" );
		WSP_CUDA( cudaSetDevice( c.e->device ) );
		WSP_CUDA( cudaEventRecord( c.ev[ 0 ], c.stream ) );
		WSP_CHECK( encodeAsync( c, offsets, batch ) );
		WSP_CUDA( cudaEventRecord( c.ev[ 1 ], c.stream ) );
		WSP_CUDA( cudaStreamSynchronize( c.stream ) );
		float ms = 0;
		cudaEventElapsedTime( &ms, c.ev[ 0 ], c.ev[ 1 ] );
		c.ms[ 1 ] += ms; c.calls[ 1 ]++;
		return WSP_OK;
	}

	// ===============================================================================================================
	// decoder
	// ===============================================================================================================
	namespace
	{
		struct ProfScope
		{
			Context& c;
			bool active;
			ProfScope( Context& ctx, int kind ) : c( ctx ), active( ctx.prof.on )
			{
				if( !active ) return;
				KernelProfile& p = c.prof;
				if( p.used + 2 > p.pool.size() )
				{
					for( int i = 0; i < 2; i++ ) { cudaEvent_t e; cudaEventCreate( &e ); p.pool.push_back( e ); }
				}
				p.kinds.push_back( kind );
				cudaEventRecord( p.pool[ p.used ], c.stream );
			}
			~ProfScope()
			{
				if( !active ) return;
				cudaEventRecord( c.prof.pool[ c.prof.used + 1 ], c.stream );
				c.prof.used += 2;
			}
		};
	}
#define WSP_KERNEL( kind, expr ) do { ProfScope _ps( c, kind ); WSP_CUDA( expr ); } while( 0 )

	// Enqueue one decoder pass for `batch` chunks x N tokens.  Reads tokens from c.tokensDev, n_past / sampling flags from device
	// scalars (so the N = 1 instance can be captured once as a CUDA graph and replayed for every step).
	static int decodeEnqueue( Context& c, int N, int batch, bool allLogits, bool sample, int* launchesOut )
	{
		Engine& e = *c.e;
		const HParams& hp = e.hp;
		const int d = hp.n_text_state, H = hp.n_text_head, T = hp.n_audio_ctx, nCtx = hp.n_text_ctx;
		cudaStream_t s = c.stream;
		const int cols = batch * N;
		const float qkScale = (float)pow( 64.0, -0.25 );   // whisper.cpp:1588, 1595, 1700
		int n = 0;
		if( N == 1 && !allLogits && c.stepMode == 2 && kern::flowSupported( d, batch, T, H, nCtx, c.refThreads, e.numSMs ) )
		{
			// steady state: ONE dataflow kernel for embedding + all layers + logits (decode_flow.cu), then the sampler
			kern::FlowArgs fa;
			fa.layers = c.flowLayers; fa.L = hp.n_text_layer; fa.B = batch; fa.maxB = c.maxB; fa.H = H; fa.nTextCtx = nCtx; fa.T = T; fa.nVocab = hp.n_vocab;
			fa.refThreads = c.refThreads;
			fa.tokEmb = e.tokEmb; fa.decPos = e.decPos; fa.lnfg = e.decLn.g; fa.lnfb = e.decLn.b;
			fa.tokens = c.tokensDev; fa.dNPast = c.dNPast;
			fa.exch = c.flowExch; fa.ctrl = c.flowCtrl; fa.logits = c.logits; fa.timing = c.stepTiming ? c.megaTiming : nullptr;
			fa.g = c.flowGeom;
			fa.timingCta = c.stepTimingCta;
			WSP_KERNEL( KK_SKINNY, kern::decodeStepFlow( fa, d, e.numSMs, s ) ); n++;
			if( sample )
			{
				kern::SampleArgs sa;
				sa.logits = c.logits; sa.probs = c.probs; sa.tieScratch = c.tieScratch; sa.B = batch; sa.nVocab = hp.n_vocab;
				sa.tokenBeg = e.tokBeg; sa.tokenSot = e.tokSot; sa.tokenSolm = e.tokSolm; sa.tokenNot = e.tokNot;
				sa.dForceTs = c.dFlags; sa.out = c.sampled; sa.nextTokens = c.tokensDev; sa.dNPast = c.dNPast; sa.N = N;
				sa.history = c.history; sa.histCap = c.histCap; sa.dStep = c.dStep;
				WSP_KERNEL( KK_OTHER, kern::sampleGreedy( sa, s ) ); n += 2;
			}
			if( launchesOut ) *launchesOut = n;
			return WSP_OK;
		}
		if( N == 1 && !allLogits && c.stepMode == 1 && kern::megaSupported( d, batch, T ) )
		{
			// steady state: one persistent kernel for embedding + all layers + logits (decode_mega.cu), then the sampler
			kern::MegaArgs ma;
			ma.layers = c.megaLayers; ma.L = hp.n_text_layer; ma.B = batch; ma.H = H; ma.nTextCtx = nCtx; ma.T = T; ma.nVocab = hp.n_vocab;
			ma.refThreads = c.refThreads;
			ma.tokEmb = e.tokEmb; ma.decPos = e.decPos; ma.lnfg = e.decLn.g; ma.lnfb = e.decLn.b;
			ma.tokens = c.tokensDev; ma.dNPast = c.dNPast;
			ma.x = c.xd; ma.q = c.qd; ma.attn = c.attnD; ma.h = c.hD; ma.logits = c.logits; ma.barrier = c.megaBarrier; ma.timing = c.stepTiming ? c.megaTiming : nullptr;
			WSP_KERNEL( KK_SKINNY, kern::decodeStepMega( ma, d, e.numSMs, s ) ); n++;
			if( sample )
			{
				kern::SampleArgs sa;
				sa.logits = c.logits; sa.probs = c.probs; sa.tieScratch = c.tieScratch; sa.B = batch; sa.nVocab = hp.n_vocab;
				sa.tokenBeg = e.tokBeg; sa.tokenSot = e.tokSot; sa.tokenSolm = e.tokSolm; sa.tokenNot = e.tokNot;
				sa.dForceTs = c.dFlags; sa.out = c.sampled; sa.nextTokens = c.tokensDev; sa.dNPast = c.dNPast; sa.N = N;
				sa.history = c.history; sa.histCap = c.histCap; sa.dStep = c.dStep;
				WSP_KERNEL( KK_OTHER, kern::sampleGreedy( sa, s ) ); n += 2;
			}
			if( launchesOut ) *launchesOut = n;
			return WSP_OK;
		}
		WSP_KERNEL( KK_OTHER, kern::embedTokens( e.tokEmb, e.decPos, c.tokensDev, c.dNPast, c.xd, batch, N, d, s ) ); n++;
		for( int il = 0; il < hp.n_text_layer; il++ )
		{
			const DecLayerW& W = e.dec[ il ];
			__half* kc = c.selfK + (size_t)il * c.maxB * nCtx * d;
			__half* vc = c.selfV + (size_t)il * c.maxB * nCtx * d;
			kern::SkinnyArgs a;
			// LN + (Q | K | V); K, V appended to the f16 self-KV cache in the epilogue (a14)
			a = kern::SkinnyArgs();
			a.W = W.wqkv; a.nOut = 3 * d; a.K = d; a.xF32 = c.xd; a.xStride = d; a.gamma = W.ln1.g; a.beta = W.ln1.b; a.nCols = cols;
			a.epi = kern::SK_QKV; a.bias = W.bqkv; a.scale = qkScale; a.outF32 = c.qd; a.ld = d; a.kCache = kc; a.vCache = vc;
			a.d = d; a.N = N; a.nTextCtx = nCtx; a.dNPast = c.dNPast;
			WSP_KERNEL( KK_SKINNY, kern::skinnyGemm( a, s ) );
			WSP_KERNEL( KK_SELF, kern::selfAttnDecode( c.qd, kc, vc, c.attnD, batch, N, H, d, nCtx, c.dNPast, c.refThreads, s ) );
			a = kern::SkinnyArgs();
			a.W = W.wo; a.nOut = d; a.K = d; a.xF16 = c.attnD; a.xStride = d; a.nCols = cols;
			a.epi = kern::SK_BIAS_RESID; a.bias = W.bo; a.outF32 = c.xd; a.ld = d;
			WSP_KERNEL( KK_SKINNY, kern::skinnyGemm( a, s ) );
			// cross attention (a15)
			a = kern::SkinnyArgs();
			a.W = W.wcq; a.nOut = d; a.K = d; a.xF32 = c.xd; a.xStride = d; a.gamma = W.lnc.g; a.beta = W.lnc.b; a.nCols = cols;
			a.epi = kern::SK_Q_SCALE; a.bias = W.bcq; a.scale = qkScale; a.outF32 = c.qd; a.ld = d;
			WSP_KERNEL( KK_SKINNY, kern::skinnyGemm( a, s ) );
			const size_t crossOff = (size_t)il * c.maxB * T * d;
			WSP_KERNEL( KK_CROSS, kern::crossAttnDecode( c.qd, c.crossK + crossOff, c.crossV + crossOff, c.attnD, batch, N, H, d, T, c.refThreads, s ) );
			a = kern::SkinnyArgs();
			a.W = W.wco; a.nOut = d; a.K = d; a.xF16 = c.attnD; a.xStride = d; a.nCols = cols;
			a.epi = kern::SK_BIAS_RESID; a.bias = W.bco; a.outF32 = c.xd; a.ld = d;
			WSP_KERNEL( KK_SKINNY, kern::skinnyGemm( a, s ) );
			// MLP (a16)
			a = kern::SkinnyArgs();
			a.W = W.w1; a.nOut = 4 * d; a.K = d; a.xF32 = c.xd; a.xStride = d; a.gamma = W.ln3.g; a.beta = W.ln3.b; a.nCols = cols;
			a.epi = kern::SK_GELU_F16; a.bias = W.b1; a.outF16 = c.hD; a.ld = 4 * d;
			WSP_KERNEL( KK_SKINNY, kern::skinnyGemm( a, s ) );
			a = kern::SkinnyArgs();
			a.W = W.w2; a.nOut = d; a.K = 4 * d; a.xF16 = c.hD; a.xStride = 4 * d; a.nCols = cols;
			a.epi = kern::SK_BIAS_RESID; a.bias = W.b2; a.outF32 = c.xd; a.ld = d;
			WSP_KERNEL( KK_SKINNY, kern::skinnyGemm( a, s ) );
			n += 8;
		}
		// final LN + logits = tok_emb^T x (a17): only the last token of each chunk unless all logits were requested
		{
			kern::SkinnyArgs a;
			a.W = e.tokEmb; a.nOut = hp.n_vocab; a.K = d; a.gamma = e.decLn.g; a.beta = e.decLn.b;
			if( allLogits ) { a.xF32 = c.xd; a.xStride = d; a.nCols = cols; }
			else { a.xF32 = c.xd + (size_t)( N - 1 ) * d; a.xStride = (int64_t)N * d; a.nCols = batch; }
			a.epi = kern::SK_LOGITS; a.outF32 = c.logits; a.ld = hp.n_vocab;
			WSP_KERNEL( KK_SKINNY, kern::skinnyGemm( a, s ) ); n++;
		}
		if( sample && !allLogits )
		{
			kern::SampleArgs sa;
			sa.logits = c.logits; sa.probs = c.probs; sa.tieScratch = c.tieScratch; sa.B = batch; sa.nVocab = hp.n_vocab;
			sa.tokenBeg = e.tokBeg; sa.tokenSot = e.tokSot; sa.tokenSolm = e.tokSolm; sa.tokenNot = e.tokNot;
			sa.dForceTs = c.dFlags; sa.out = c.sampled; sa.nextTokens = c.tokensDev; sa.dNPast = c.dNPast; sa.N = N;
			sa.history = c.history; sa.histCap = c.histCap; sa.dStep = c.dStep;
			WSP_KERNEL( KK_OTHER, kern::sampleGreedy( sa, s ) ); n += 2;
		}
		if( launchesOut ) *launchesOut = n;
		return WSP_OK;
	}

	// Submit one decoder pass: the captured CUDA graph for the N = 1 steady state, plain launches otherwise.
	static int decodeSubmit( Context& c, int N, int batch, bool allLogits, bool sample )
	{
		int n = 0;
		if( N == 1 && !allLogits && sample && c.useGraph && !c.prof.on )
		{
			if( !c.stepGraph || c.stepGraphBatch != batch )
			{
				if( c.stepGraph ) { cudaGraphExecDestroy( c.stepGraph ); c.stepGraph = nullptr; }
				cudaGraph_t graph = nullptr;
				WSP_CUDA( cudaStreamBeginCapture( c.stream, cudaStreamCaptureModeThreadLocal ) );
				const int rc = decodeEnqueue( c, 1, batch, false, true, &n );
				cudaError_t ce = cudaStreamEndCapture( c.stream, &graph );
				if( rc < 0 ) { if( graph ) cudaGraphDestroy( graph ); return rc; }
				WSP_CUDA( ce );
				ce = cudaGraphInstantiate( &c.stepGraph, graph, 0 );
				cudaGraphDestroy( graph );
				WSP_CUDA( ce );
				c.stepGraphBatch = batch;
				c.stepGraphLaunches = n;
			}
			WSP_CUDA( cudaGraphLaunch( c.stepGraph, c.stream ) );
			g_launchCount.fetch_add( (uint64_t)c.stepGraphLaunches, std::memory_order_relaxed );
			return WSP_OK;
		}
		WSP_CHECK( decodeEnqueue( c, N, batch, allLogits, sample, &n ) );
		launched( n );
		if( allLogits )
		{
			// probabilities for every row (the oracle keeps logits and probs of all N tokens, whisper.cpp:1855-1859); no sampling
			WSP_CUDA( kern::softmaxRows( c.logits, c.probs, batch * N, c.e->hp.n_vocab, c.stream ) );
			launched();
		}
		return WSP_OK;
	}

	int ctxDecode( Context& c, const int32_t* tokensHost, int nTokens, int nPast, int batch, uint32_t flags, wsp_token_data* sampledHost )
	{
		const HParams& hp = c.e->hp;
		const bool devTokens = ( flags & WSP_DECODE_DEVICE_TOKENS ) != 0;
		const bool allLogits = ( flags & WSP_DECODE_ALL_LOGITS ) != 0;
		const bool sample = ( flags & WSP_DECODE_NO_SAMPLE ) == 0;
		if( batch < 1 || batch > c.maxB ) return fail( WSP_E_BOUNDS, "batch exceeds the context's max_batch" );
		if( nTokens < 1 || nTokens > kMaxDecodeTokens ) return fail( WSP_E_BOUNDS, "n_tokens out of range" );
		if( allLogits && nTokens > kAllLogitsTokens ) return fail( WSP_E_BOUNDS, "ALL_LOGITS supports at most 8 tokens" );
		if( devTokens && nTokens != 1 ) return fail( WSP_E_INVALIDARG, "DEVICE_TOKENS implies n_tokens == 1" );
		if( !devTokens && !tokensHost ) return fail( WSP_E_POINTER, "tokens" );
		if( !devTokens && ( nPast < 0 || nPast + nTokens > hp.n_text_ctx ) ) return fail( WSP_E_BOUNDS, "n_past + n_tokens exceeds n_text_ctx" );
		// device-fed tokens continue from the device's n_past: its host mirror guards the KV cache and the positional table
		if( devTokens && c.nPastHost + nTokens > hp.n_text_ctx ) return fail( WSP_E_BOUNDS, "n_past + n_tokens exceeds n_text_ctx" );
		WSP_CUDA( cudaSetDevice( c.e->device ) );
		cudaStream_t s = c.stream;
		WSP_CUDA( cudaEventRecord( c.ev[ 0 ], s ) );
		if( !devTokens )
		{
			for( int i = 0; i < batch * nTokens; i++ )
				if( tokensHost[ i ] < 0 || tokensHost[ i ] >= hp.n_vocab ) return fail( WSP_E_INVALIDARG, "token id out of range" );
			WSP_CUDA( cudaMemcpyAsync( c.tokensDev, tokensHost, (size_t)batch * nTokens * 4, cudaMemcpyHostToDevice, s ) );
			WSP_CUDA( kern::setInts( c.dNPast, nPast, 0, s ) );
			WSP_CUDA( kern::setInts( c.dFlags, ( flags & WSP_DECODE_FORCE_TIMESTAMP ) ? 1 : 0, ( flags & WSP_DECODE_INITIAL ) ? 1 : 0, s ) );
			launched( 2 );
		}
		WSP_CHECK( decodeSubmit( c, nTokens, batch, allLogits, sample ) );
		// the device advances n_past only when it samples (advance_kernel): mirror exactly that
		if( !devTokens ) c.nPastHost = nPast;
		if( sample && !allLogits ) c.nPastHost += nTokens;
		c.lastLogitRows = allLogits ? batch * nTokens : batch;
		WSP_CUDA( cudaEventRecord( c.ev[ 1 ], s ) );
		if( sampledHost && sample && !allLogits )
			WSP_CUDA( cudaMemcpyAsync( sampledHost, c.sampled, (size_t)batch * sizeof( wsp_token_data ), cudaMemcpyDeviceToHost, s ) );
		WSP_CUDA( cudaStreamSynchronize( s ) );
		float ms = 0;
		cudaEventElapsedTime( &ms, c.ev[ 0 ], c.ev[ 1 ] );
		c.ms[ 2 ] += ms; c.calls[ 2 ]++;
		return WSP_OK;
	}

	int ctxUploadPcm( Context& c, int slot, const float* pcmHost, int nSamples )
	{
		if( slot < 0 || slot >= c.maxB ) return fail( WSP_E_BOUNDS, "chunk slot out of range" );
		if( ( !pcmHost && nSamples != 0 ) || nSamples < 0 ) return fail( WSP_E_INVALIDARG, "pcm" );
		WSP_CUDA( cudaSetDevice( c.e->device ) );
		MelSlot& s = c.slots[ slot ];
		if( nSamples > s.pcmCap )
		{
			if( s.pcm ) { WSP_CUDA( cudaStreamSynchronize( c.stream ) ); cudaFree( s.pcm ); s.pcm = nullptr; }
			WSP_CHECK( devAlloc( s.pcm, (size_t)( nSamples > 0 ? nSamples : 1 ) ) );
			s.pcmCap = nSamples;
		}
		s.pcmSamples = nSamples;
		WSP_CUDA( cudaMemcpyAsync( s.pcm, pcmHost, (size_t)nSamples * 4, cudaMemcpyHostToDevice, c.stream ) );
		WSP_CUDA( cudaStreamSynchronize( c.stream ) );
		return WSP_OK;
	}

	// Instrumented decoder pass: nSteps single-token steps launched kernel by kernel with a CUDA event pair around every launch
	// (on the launching stream).  Continues from the context's current decoder state; used by bench.py for the roofline figure.
	int ctxProfileDecode( Context& c, int batch, int nSteps, float* msByKind, int* launchesByKind )
	{
		if( batch < 1 || batch > c.maxB || nSteps < 1 ) return fail( WSP_E_INVALIDARG, "batch / steps" );
		if( c.nPastHost + nSteps > c.e->hp.n_text_ctx ) return fail( WSP_E_BOUNDS, "n_past + n_steps exceeds n_text_ctx" );
		WSP_CUDA( cudaSetDevice( c.e->device ) );
		c.prof.on = true;
		c.prof.used = 0;
		c.prof.kinds.clear();
		int rc = WSP_OK;
		for( int i = 0; i < nSteps && rc >= 0; i++ ) { rc = decodeSubmit( c, 1, batch, false, true ); if( rc >= 0 ) c.nPastHost++; }
		c.prof.on = false;
		if( rc < 0 ) return rc;
		WSP_CUDA( cudaStreamSynchronize( c.stream ) );
		for( int k = 0; k < KK_COUNT; k++ ) { msByKind[ k ] = 0; launchesByKind[ k ] = 0; }
		for( size_t i = 0; i < c.prof.kinds.size(); i++ )
		{
			float ms = 0;
			cudaEventElapsedTime( &ms, c.prof.pool[ 2 * i ], c.prof.pool[ 2 * i + 1 ] );
			msByKind[ c.prof.kinds[ i ] ] += ms;
			launchesByKind[ c.prof.kinds[ i ] ]++;
		}
		return WSP_OK;
	}

	// ===============================================================================================================
	// the measured path: mel + encode + n_decode greedy steps, tokens fed back on the device
	// ===============================================================================================================
	int ctxRunChunks( Context& c, const float* const* pcm, const int32_t* nSamples, int batch, const int32_t* prompt, int nPrompt, int nDecode,
		int32_t* tokensOut, float* stageMs, bool resident )
	{
		const HParams& hp = c.e->hp;
		if( batch < 1 || batch > c.maxB ) return fail( WSP_E_BOUNDS, "batch exceeds the context's max_batch" );
		if( !prompt || nPrompt < 1 || nPrompt > kMaxDecodeTokens ) return fail( WSP_E_INVALIDARG, "prompt" );
		if( nDecode < 1 || nPrompt + nDecode > hp.n_text_ctx || nDecode > c.histCap ) return fail( WSP_E_BOUNDS, "n_prompt + n_decode exceeds n_text_ctx" );
		if( !resident && ( !pcm || !nSamples ) ) return fail( WSP_E_POINTER, "pcm" );
		WSP_CUDA( cudaSetDevice( c.e->device ) );
		cudaStream_t s = c.stream;
		WSP_CUDA( cudaEventRecord( c.ev[ 0 ], s ) );
		if( resident )
		{
			// inputs already in HBM (wsp_upload_pcm): the log-mel front end still runs inside the timed region
			for( int b = 0; b < batch; b++ )
				if( !c.slots[ b ].pcm ) return fail( WSP_E_INVALIDARG, "no resident PCM in slot " + std::to_string( b ) + " (call wsp_upload_pcm)" );
			WSP_CHECK( melBatchFromDevicePcm( c, batch, [ & ]( int b ) { return (const float*)c.slots[ b ].pcm; }, [ & ]( int b ) { return c.slots[ b ].pcmSamples; } ) );
		}
		else
		{
			size_t total = 0;
			for( int b = 0; b < batch; b++ ) { if( nSamples[ b ] < 0 ) return fail( WSP_E_INVALIDARG, "n_samples" ); total += (size_t)nSamples[ b ]; }
			WSP_CHECK( ensurePcm( c, total ? total : 1 ) );
			std::vector<size_t> offs( (size_t)batch );
			size_t off = 0;
			for( int b = 0; b < batch; b++ )
			{
				if( nSamples[ b ] > 0 )
					WSP_CUDA( cudaMemcpyAsync( c.pcmDev + off, pcm[ b ], (size_t)nSamples[ b ] * 4, cudaMemcpyHostToDevice, s ) );
				offs[ (size_t)b ] = off;
				off += (size_t)nSamples[ b ];
			}
			WSP_CHECK( melBatchFromDevicePcm( c, batch, [ & ]( int b ) { return (const float*)( c.pcmDev + offs[ (size_t)b ] ); }, [ & ]( int b ) { return (int)nSamples[ b ]; } ) );
		}
		WSP_CUDA( cudaEventRecord( c.ev[ 1 ], s ) );
		WSP_CHECK( encodeAsync( c, nullptr, batch ) );
		WSP_CUDA( cudaEventRecord( c.ev[ 2 ], s ) );
		// prompt step: same prompt for every chunk, first sample uses the initial-timestamp rules (whisper.cpp:2943)
		std::vector<int32_t> ptoks( (size_t)batch * nPrompt );
		for( int b = 0; b < batch; b++ ) for( int i = 0; i < nPrompt; i++ ) ptoks[ (size_t)b * nPrompt + i ] = prompt[ i ];
		WSP_CUDA( cudaMemcpyAsync( c.tokensDev, ptoks.data(), ptoks.size() * 4, cudaMemcpyHostToDevice, s ) );
		WSP_CUDA( kern::setInts( c.dNPast, 0, 0, s ) );
		WSP_CUDA( kern::setInts( c.dFlags, 1, 1, s ) );
		WSP_CUDA( kern::setInts( c.dStep, 0, 0, s ) );
		launched( 3 );
		WSP_CHECK( decodeSubmit( c, nPrompt, batch, false, true ) );
		for( int i = 1; i < nDecode; i++ )
			WSP_CHECK( decodeSubmit( c, 1, batch, false, true ) );
		c.nPastHost = nPrompt + nDecode - 1;
		WSP_CUDA( cudaEventRecord( c.ev[ 3 ], s ) );
		c.lastLogitRows = batch;
		// one D2H of the token log at the end: [batch][n_decode] int32
		std::vector<int32_t> hist( (size_t)batch * c.histCap );
		WSP_CUDA( cudaMemcpyAsync( hist.data(), c.history, hist.size() * 4, cudaMemcpyDeviceToHost, s ) );
		WSP_CUDA( cudaStreamSynchronize( s ) );
		if( tokensOut )
			for( int b = 0; b < batch; b++ )
				for( int i = 0; i < nDecode; i++ ) tokensOut[ (size_t)b * nDecode + i ] = hist[ (size_t)b * c.histCap + i ];
		float m0 = 0, m1 = 0, m2 = 0;
		cudaEventElapsedTime( &m0, c.ev[ 0 ], c.ev[ 1 ] );
		cudaEventElapsedTime( &m1, c.ev[ 1 ], c.ev[ 2 ] );
		cudaEventElapsedTime( &m2, c.ev[ 2 ], c.ev[ 3 ] );
		c.ms[ 0 ] += m0; c.ms[ 1 ] += m1; c.ms[ 2 ] += m2; c.calls[ 0 ]++; c.calls[ 1 ]++; c.calls[ 2 ] += nDecode;
		if( stageMs ) { stageMs[ 0 ] = m0; stageMs[ 1 ] = m1; stageMs[ 2 ] = m2; }
		return WSP_OK;
	}
}
