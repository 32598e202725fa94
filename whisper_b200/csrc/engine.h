// Engine (weights on one B200) and Context (per-stream state for up to maxBatch independent 30 s chunks).
//
// Reference counterparts: ModelBuffers (Whisper/Whisper/ModelBuffers.h:8-111) / WhisperModel (loader), and WhisperContext
// (Whisper/Whisper/WhisperContext.cpp:138-639 — the network definition) with its KeyValueBuffers (KeyValueBuffers.h:7-53).
// Oracle: whisper_encode / whisper_decode (Whisper/source/whisper.cpp:1084-1872).
#pragma once
#include "../../include/whisper_b200.h"
#include "attn_enc.cuh"
#include "decode_flow.cuh"
#include "decode_mega.cuh"
#include "gemm_tc.cuh"
#include "kernels.cuh"
#include "model.h"
#include <atomic>
#include <string>
#include <vector>

namespace wsp
{
	extern thread_local std::string g_lastError;
	extern std::atomic<uint64_t> g_launchCount;
	int fail( int status, const std::string& what );
	int cudaFail( cudaError_t e, const char* what );

#define WSP_CUDA( expr )                                                        \
	do {                                                                        \
		cudaError_t _e = ( expr );                                              \
		if( _e != cudaSuccess ) return ::wsp::cudaFail( _e, #expr );            \
	} while( 0 )
#define WSP_CHECK( expr )                                                       \
	do {                                                                        \
		int _s = ( expr );                                                      \
		if( _s < 0 ) return _s;                                                 \
	} while( 0 )

	constexpr int kFrames = 3000;          // mel frames per 30 s window
	constexpr int kFramesPad = 3002;       // + one zero halo row each side (conv padding = 1)
	constexpr int kConv1KTap = 128;        // 80 input channels padded to two 64-wide K blocks per tap
	constexpr int kMaxDecodeTokens = 256;  // tokens per chunk in one decoder call (prompt <= 224 + 4)
	constexpr int kAllLogitsTokens = 8;

	struct LnW { float* g = nullptr; float* b = nullptr; };
	struct EncLayerW
	{
		LnW ln1, ln2;
		__half* wqkv = nullptr; float* bqkv = nullptr;   // [3d][d], (q.b | 0 | v.b)
		__half* wo = nullptr; float* bo = nullptr;
		__half* w1 = nullptr; float* b1 = nullptr;       // [4d][d]
		__half* w2 = nullptr; float* b2 = nullptr;       // [d][4d]
	};
	struct DecLayerW
	{
		LnW ln1, lnc, ln3;
		__half* wqkv = nullptr; float* bqkv = nullptr;
		__half* wo = nullptr; float* bo = nullptr;
		__half* wcq = nullptr; float* bcq = nullptr;
		__half* wco = nullptr; float* bco = nullptr;
		__half* w1 = nullptr; float* b1 = nullptr;
		__half* w2 = nullptr; float* b2 = nullptr;
	};

	struct Engine
	{
		int device = 0;
		int numSMs = 148;
		HParams hp{};
		int tokEot = 0, tokSot = 0, tokPrev = 0, tokSolm = 0, tokNot = 0, tokBeg = 0;
		uint8_t* arena = nullptr;
		size_t arenaSize = 0, arenaUsed = 0;

		__half* conv1w = nullptr; float* conv1b = nullptr;   // [d][3*128]
		__half* conv2w = nullptr; float* conv2b = nullptr;   // [d][3*d]
		float* encPos = nullptr;                             // [1500][d]
		LnW encLnPost;
		std::vector<EncLayerW> enc;
		__half* crossW = nullptr; float* crossB = nullptr;   // [L*2d][d] (K_l | V_l), bias (0 | v.b)
		float* decPos = nullptr;                             // [n_text_ctx][d]
		__half* tokEmb = nullptr;                            // [n_vocab][d]
		LnW decLn;
		std::vector<DecLayerW> dec;
		kern::MelTables mel;

		~Engine();
	};

	int createEngine( const ModelFile& m, int device, const void* devImage, uint64_t imageSize, Engine** out );

	struct MelSlot
	{
		float* mel = nullptr;   // [80][nLen]
		int cap = 0;            // frames allocated
		int nLen = 0;
		float* pcm = nullptr;   // optional device-resident PCM (wsp_upload_pcm), for the HBM-resident bench leg
		int pcmCap = 0;
		int pcmSamples = 0;
	};

	// per-kernel-kind device time of the decoder, collected by an instrumented (un-graphed) pass: wsp_profile_decode
	enum KernelKind { KK_SKINNY = 0, KK_CROSS = 1, KK_SELF = 2, KK_OTHER = 3, KK_COUNT = 4 };
	struct KernelProfile
	{
		bool on = false;
		std::vector<cudaEvent_t> pool;
		std::vector<int> kinds;     // kind of interval i = [pool[2i], pool[2i+1]]
		size_t used = 0;
	};

	struct Context
	{
		Engine* e = nullptr;
		int maxB = 0;
		cudaStream_t stream = nullptr;
		std::vector<MelSlot> slots;
		int* melMax = nullptr;          // [maxB] ordered-int maxima
		float* pcmDev = nullptr; size_t pcmCap = 0;
		uint64_t devBytes = 0;          // device memory this context holds (timingsPrint's memory table)

		// encoder workspaces
		__half* melF16 = nullptr;       // [maxB][3002][80]
		__half* conv1 = nullptr;        // [maxB][3002][d]
		float* x = nullptr;             // [maxB*1500][d] residual stream
		__half* xn = nullptr;           // [maxB*1500][d]
		__half* q = nullptr;            // [maxB][H][1500][64]
		__half* k = nullptr;
		__half* vt = nullptr;           // [maxB][H][64][Tp]
		__half* attn = nullptr;         // [maxB*1500][d]
		__half* h = nullptr;            // [maxB*1500][4d]
		__half* crossK = nullptr;       // [L][maxB][H][1500][64]
		__half* crossV = nullptr;
		int Tp = 0;

		// decoder state
		__half* selfK = nullptr;        // [L][maxB][H][n_text_ctx][64] (head-major rows)
		__half* selfV = nullptr;
		float* xd = nullptr;            // [maxB*kMaxDecodeTokens][d]
		float* qd = nullptr;
		__half* attnD = nullptr;
		__half* hD = nullptr;           // [maxB*kMaxDecodeTokens][4d]
		float* logits = nullptr;        // [maxB*kAllLogitsTokens][n_vocab]
		float* probs = nullptr;
		int* tieScratch = nullptr;      // [maxB][n_vocab + 1024]: survivor lists of the sampler's exact tie emulation
		int* tokensDev = nullptr;       // [maxB*kMaxDecodeTokens]
		int* dNPast = nullptr;          // device scalars: n_past | flags[2] | step
		int* dFlags = nullptr;
		int* dStep = nullptr;
		kern::TokenData* sampled = nullptr;   // [maxB]
		int* history = nullptr;         // [maxB][histCap]
		int histCap = 0;
		int lastLogitRows = 0;
		int nPastHost = 0;              // host mirror of the device-resident n_past (it advances on the device between graph replays)
		int debugEncLayers = -1;
		int refThreads = 4;             // reference CPU thread count whose V^T*P arithmetic the decoder reproduces (0 = exact)

		// tensor maps (built once; M of a launch limits the rows touched)
		CUtensorMap mapMel, mapConv1Even, mapConv1Odd, mapXn, mapAttn, mapH;
		CUtensorMap mapConv1W, mapConv2W, mapCrossW;
		std::vector<CUtensorMap> mapWqkv, mapWo, mapW1, mapW2;
		CUtensorMap mapQ, mapK, mapVt;
		int bnD = 128, bn3D = 128, bn4D = 128, bnCross = 128;

		// persistent decoder-step kernel (decode_mega.cu): per-layer pointer table on the device, grid-barrier counter
		kern::MegaLayer* megaLayers = nullptr;
		unsigned* megaBarrier = nullptr;
		unsigned long long* megaTiming = nullptr;   // [4096] debug marks
		// dataflow decoder-step kernel (decode_flow.cu): per-layer pointer table, per-CTA bias slabs, exchange buffers, launch epoch
		kern::FlowLayer* flowLayers = nullptr;
		float* flowBias = nullptr;
		uint8_t* flowExch = nullptr;
		unsigned* flowCtrl = nullptr;
		kern::FlowGeom flowGeom;
		// N = 1 decoder step: 2 = dataflow kernel (default), 1 = round 1's barrier kernel, 0 = one kernel per op
		int stepMode = 2;
		int stepTimingCta = 0;
		bool stepTiming = false;        // record %globaltimer marks of CTA 0 into megaTiming (wsp_debug_step_timing)

		// decode CUDA graph (N = 1 steady state)
		cudaGraphExec_t stepGraph = nullptr;
		int stepGraphBatch = 0;
		int stepGraphLaunches = 0;
		bool useGraph = true;

		KernelProfile prof;
		cudaEvent_t timerEv[ 2 ] = { nullptr, nullptr };

		// timing
		cudaEvent_t ev[ 4 ] = { nullptr, nullptr, nullptr, nullptr };
		float ms[ 4 ] = { 0, 0, 0, 0 };
		int calls[ 4 ] = { 0, 0, 0, 0 };

		~Context();
	};

	int createContext( Engine* e, int maxBatch, Context** out );
	int ctxPcmToMel( Context& c, int slot, const float* pcmHost, int nSamples );
	int ctxPcmToMelWindow( Context& c, int slot, const float* pcmHost, int nSamples, int nFrames, const float* forcedMax, float* maxOut );
	int ctxSetMel( Context& c, int slot, const float* melHost, int nLen );
	int ctxEncode( Context& c, const int32_t* offsets, int batch );
	int ctxDecode( Context& c, const int32_t* tokensHost, int nTokens, int nPast, int batch, uint32_t flags, wsp_token_data* sampledHost );
	int ctxUploadPcm( Context& c, int slot, const float* pcmHost, int nSamples );
	int ctxProfileDecode( Context& c, int batch, int nSteps, float* msByKind, int* launchesByKind );
	int ctxRunChunks( Context& c, const float* const* pcm, const int32_t* nSamples, int batch, const int32_t* prompt, int nPrompt, int nDecode, int32_t* tokensOut, float* stageMs, bool resident );
}
