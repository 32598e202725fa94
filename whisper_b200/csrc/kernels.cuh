// Non-GEMM kernels of the hot path (HBM-bound): log-mel front end, LayerNorm, encoder window gather, token embedding,
// and the decoder's per-token kernels (skinny weight-streaming GEMM, self/cross attention over the f16 KV memories, sampler).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace kern
{
	// Launch with the programmatic-stream-serialization attribute (PDL).  Kernels launched this way call pdl_launch_dependents() first
	// thing and pdl_wait() before touching anything a predecessor wrote.  WSP_PDL=0 in the environment disables the attribute.
	bool pdlEnabled();
	template<class... KArgs, class... Args>
	inline cudaError_t launchPdl( void ( *kernel )( KArgs... ), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args )
	{
		cudaLaunchConfig_t cfg{};
		cfg.gridDim = grid;
		cfg.blockDim = block;
		cfg.dynamicSmemBytes = smem;
		cfg.stream = s;
		cudaLaunchAttribute at[ 1 ];
		at[ 0 ].id = cudaLaunchAttributeProgrammaticStreamSerialization;
		at[ 0 ].val.programmaticStreamSerializationAllowed = 1;
		cfg.attrs = at;
		cfg.numAttrs = pdlEnabled() ? 1 : 0;
		return cudaLaunchKernelEx( &cfg, kernel, static_cast<KArgs>( args )... );
	}

	// ---- log-mel (a1: whisper.cpp:2060-2181) -------------------------------------------------------------------
	struct MelTables
	{
		const float* hann = nullptr;     // [400]  f32( 0.5*(1-cos(2*pi*i/400)) )
		const double* cosT = nullptr;    // [400]  cos(2*pi*m/400)
		const double* sinT = nullptr;    // [400]
		const float* filters = nullptr;  // [80][201] from the model file
		const double2* twiddle = nullptr; // [400]  { cosT, sinT } interleaved: one 16-byte shared-memory read per DFT term
		const short2* band = nullptr;     // [80]   [x, y) = the bins where the band's filter row is non-zero (real filter banks are triangular)
	};
	// pcm [nSamples] -> raw log10 mel [80][nLen] (row = band) and the running maximum (ordered-int encoded float) in *maxSlot
	// up to MEL_BATCH chunks transformed and normalised by one launch each (power, normalise); maxSlots[b] is chunk b's maximum
	constexpr int MEL_BATCH = 16;
	struct MelBatch
	{
		const float* pcm[ MEL_BATCH ];
		float* mel[ MEL_BATCH ];       // [80][nLen[b]]
		int nSamples[ MEL_BATCH ];
		int nLen[ MEL_BATCH ];
		int* maxSlots;                 // [count] consecutive
		int count;
	};
	cudaError_t melBatch( const MelTables& tb, const MelBatch& mb, cudaStream_t s );
	cudaError_t melPower( const MelTables& tb, const float* pcm, int nSamples, int nLen, float* melRaw, int* maxSlot, cudaStream_t s );
	// in place: clamp to (max - 8), (x + 4) / 4
	cudaError_t melNormalize( float* mel, int count, const int* maxSlot, cudaStream_t s );
	// mel f32 [80][nLen] at frame offset -> time-major f16 window [1 + 3000 + 1][80] with zero halo rows (conv padding)
	cudaError_t melWindow( const float* mel, int nLen, int offset, __half* dst, int frames, cudaStream_t s );

	// ---- LayerNorm + affine -> f16 (a5, a11) -------------------------------------------------------------------------
	cudaError_t layerNormF16( const float* x, const float* gamma, const float* beta, __half* out, int rows, int d, cudaStream_t s );

	// ---- decoder -------------------------------------------------------------------------------------------------------
	// x[b*N+i][:] = f32(te[token[b*N+i]][:]) + pe[nPast+i][:]     (a13)
	cudaError_t embedTokens( const __half* te, const float* pe, const int* tokens, const int* dNPast, float* x, int B, int N, int d, cudaStream_t s );

	enum SkinnyEpi : int
	{
		SK_QKV = 0,        // rows [0,d): q f32 = (acc+b)*s ; [d,2d): K cache f16 = acc*s ; [2d,3d): V cache f16 = acc+b
		SK_BIAS_RESID = 1, // x f32 += acc + b
		SK_Q_SCALE = 2,    // q f32 = (acc+b)*s
		SK_GELU_F16 = 3,   // h f16 = gelu(acc+b)
		SK_LOGITS = 4,     // logits f32 = acc
	};
	struct SkinnyArgs
	{
		const __half* W = nullptr;     // [nOut][K] f16 row-major
		int nOut = 0, K = 0;
		// input columns: either f32 rows + LayerNorm (gamma != null) or f16 rows
		const float* xF32 = nullptr;
		const __half* xF16 = nullptr;
		int64_t xStride = 0;           // elements between consecutive columns
		const float* gamma = nullptr;
		const float* beta = nullptr;
		int nCols = 0;
		// epilogue
		int epi = 0;
		const float* bias = nullptr;
		float scale = 1.0f;
		float* outF32 = nullptr;       // q / x / logits  [col][ld]
		__half* outF16 = nullptr;      // h             [col][ld]
		int ld = 0;
		__half* kCache = nullptr;      // [B][H][nTextCtx][64] for this layer (head-major)
		__half* vCache = nullptr;
		int d = 0, N = 0, nTextCtx = 0;
		const int* dNPast = nullptr;
	};
	cudaError_t skinnyGemm( const SkinnyArgs& a, cudaStream_t s );

	// refThreads: reproduce the reference's f16-accumulated V^T*P for that many CPU threads (see kernels_decode.cu); 0 = exact f32
	// self attention over the f16 self-KV cache, causal: query i of chunk b sees keys [0, nPast+i]   (a14)
	cudaError_t selfAttnDecode( const float* q, const __half* kCache, const __half* vCache, __half* out, int B, int N, int H, int d, int nTextCtx, const int* dNPast, int refThreads, cudaStream_t s );
	// cross attention over the f16 cross-KV memory [b][h][T][64]                                       (a15)
	cudaError_t crossAttnDecode( const float* q, const __half* kMem, const __half* vMem, __half* out, int B, int N, int H, int d, int T, int refThreads, cudaStream_t s );

	// softmax + greedy sampling with the Whisper timestamp rules (a17 softmax, a18: whisper.cpp:1875-1964)
	struct TokenData
	{
		int id, tid;
		float p, pt, ptsum;
	};
	struct SampleArgs
	{
		const float* logits = nullptr;  // [B][nVocab] (last token of each chunk)
		float* probs = nullptr;         // [B][nVocab] out
		int B = 0, nVocab = 0;
		int tokenBeg = 0, tokenSot = 0, tokenSolm = 0, tokenNot = 0;
		const int* dForceTs = nullptr;  // device flags: [0] force_timestamp, [1] is_initial (so one CUDA graph serves every step)
		TokenData* out = nullptr;       // [B]
		int* nextTokens = nullptr;      // [B] device feedback for the next step (may be null)
		int* dNPast = nullptr;          // advanced by N after sampling (may be null)
		int N = 0;
		int* history = nullptr;         // optional [B][histCap] token log, written at column *dStep
		int histCap = 0;
		int* dStep = nullptr;
		int rowInSmem = 0;              // set by sampleGreedy: the row fits the 227 KB shared memory
		int* tieScratch = nullptr;      // [B][nVocab + 1024] ints: survivor lists of the exact tie emulation (null: one-thread emulation)
	};
	cudaError_t sampleGreedy( const SampleArgs& a, cudaStream_t s );

	// row-wise softmax of [rows][n] (all-logits test mode)
	cudaError_t softmaxRows( const float* logits, float* probs, int rows, int n, cudaStream_t s );
	// one-time function attributes (dynamic shared memory opt-in) so that nothing but launches happens under stream capture
	cudaError_t prepare( int maxK );

	// tiny helpers
	cudaError_t setInts( int* dst, int a, int b, cudaStream_t s );
}
