// TEST INFRASTRUCTURE — parity oracle, "reference" kind (oracle/_ref/liboracle_ref.so).
//
// This translation unit textually includes the reference's UNMODIFIED Whisper/source/whisper.cpp from
// /root/reference (the same trick Whisper/whisperCom.cpp:52 uses) so that the private state of
// whisper_context (logits, KV memories, mel) can be read back, and exports a flat C API (ora_*) for
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  Nothing in the
// product path (whisper_b200/) may link, load or call this.
//
// No reference source is copied into this repository: the #include below resolves through
// -I/root/reference/Whisper/source at build time (oracle/Makefile); only the built .so travels.
#include "whisper.cpp"

#include <chrono>
#include <mutex>

// ---------------------------------------------------------------------------------------------------
// logger sink (declared in shim/Utils/Logger.h)
static int g_logLevel = 1;   // 0 error, 1 warning, 2 info, 3 debug
static void vlog( int lvl, const char* f, va_list ap )
{
	if( lvl > g_logLevel ) return;
	vfprintf( stderr, f, ap );
	fputc( '\n', stderr );
}
extern "C" {
void logError( const char8_t* f, ... ) { va_list ap; va_start( ap, f ); vlog( 0, (const char*)f, ap ); va_end( ap ); }
void logWarning( const char8_t* f, ... ) { va_list ap; va_start( ap, f ); vlog( 1, (const char*)f, ap ); va_end( ap ); }
void logInfo( const char8_t* f, ... ) { va_list ap; va_start( ap, f ); vlog( 2, (const char*)f, ap ); va_end( ap ); }
void logDebug( const char8_t* f, ... ) { va_list ap; va_start( ap, f ); vlog( 3, (const char*)f, ap ); va_end( ap ); }
}

// ---------------------------------------------------------------------------------------------------
// trace recorder: named tensors at the reference's own trace points
namespace
{
	struct TraceItem
	{
		std::string name;
		int64_t ne[ 4 ];
		std::vector<float> data;
	};
	struct Delayed { std::string name; const ggml_tensor* t; };
	bool g_traceOn = false;
	bool g_gapLogOn = false;
	int g_gapVocab = 0;
	std::vector<float> g_gapLog;
	std::vector<TraceItem> g_trace;
	std::vector<Delayed> g_delayed;

	void record( const char* name, const ggml_tensor* t )
	{
		TraceItem it;
		it.name = name;
		for( int i = 0; i < 4; i++ ) it.ne[ i ] = t->ne[ i ];
		const size_t n = (size_t)ggml_nelements( t );
		it.data.resize( n );
		// honour strides: traced tensors may be permuted views
		const char* base = (const char*)t->data;
		size_t o = 0;
		for( int64_t i3 = 0; i3 < t->ne[ 3 ]; i3++ )
			for( int64_t i2 = 0; i2 < t->ne[ 2 ]; i2++ )
				for( int64_t i1 = 0; i1 < t->ne[ 1 ]; i1++ )
					for( int64_t i0 = 0; i0 < t->ne[ 0 ]; i0++, o++ )
					{
						const char* p = base + i0 * t->nb[ 0 ] + i1 * t->nb[ 1 ] + i2 * t->nb[ 2 ] + i3 * t->nb[ 3 ];
						if( t->type == GGML_TYPE_F32 ) it.data[ o ] = *(const float*)p;
						else if( t->type == GGML_TYPE_F16 ) it.data[ o ] = ggml_fp16_to_fp32( *(const ggml_fp16_t*)p );
						else if( t->type == GGML_TYPE_I32 ) it.data[ o ] = (float)*(const int32_t*)p;
						else it.data[ o ] = 0;
					}
		g_trace.push_back( std::move( it ) );
	}
}
namespace Tracing
{
	void tensor( const ItemName& name, const ggml_tensor* t ) { if( g_traceOn ) record( name.text, t ); }
	void delayTensor( const ItemName& name, const ggml_tensor* t ) { if( g_traceOn ) g_delayed.push_back( { name.text, t } ); }
	void writeDelayedTensors()
	{
		for( const auto& d : g_delayed ) record( d.name.c_str(), d.t );
		g_delayed.clear();
	}
	void vector( const ItemName& name, const std::vector<float>& v )
	{
		if( g_gapLogOn && 0 == strcmp( name.text, "probs" ) && g_gapVocab > 0 && v.size() >= (size_t)g_gapVocab )
		{
			// top-2 gap of the last row in logit units: log( p1 / p2 ) — how close the reference's greedy decision was
			const float* p = v.data() + ( v.size() - (size_t)g_gapVocab );
			float a = 0, b = 0;
			for( int i = 0; i < g_gapVocab; i++ )
			{
				if( p[ i ] > a ) { b = a; a = p[ i ]; }
				else if( p[ i ] > b ) b = p[ i ];
			}
			g_gapLog.push_back( b > 0 ? logf( a / b ) : 1e9f );
		}
		if( !g_traceOn ) return;
		TraceItem it;
		it.name = name.text;
		it.ne[ 0 ] = (int64_t)v.size(); it.ne[ 1 ] = it.ne[ 2 ] = it.ne[ 3 ] = 1;
		it.data = v;
		g_trace.push_back( std::move( it ) );
	}
}

// ---------------------------------------------------------------------------------------------------
// flat C API
extern "C" {

void ora_set_log_level( int lvl ) { g_logLevel = lvl; }
const char* ora_system_info() { return whisper_print_system_info(); }

whisper_context* ora_init( const char* path )
{
	whisper_context* c = whisper_init( path );
	if( c )
	{
		// whisper_context leaves these members uninitialised (whisper.cpp:425-431) and only whisper_full() sets them
		// (:2804-2807, :2836).  whisper_encode/whisper_decode read exp_n_audio_ctx (:1092, :1528), so a caller that drives
		// encode/decode directly — as this wrapper does — must give them their documented defaults ("0 - use default").
		c->exp_n_audio_ctx = 0;
		c->t_beg = 0;
		c->t_last = 0;
		c->tid_last = 0;
	}
	return c;
}
void ora_free( whisper_context* c ) { if( c ) whisper_free( c ); }

// hparams in file order (whisper.cpp:477-487)
void ora_hparams( whisper_context* c, int32_t* out11 )
{
	const auto& h = c->model.hparams;
	const int32_t v[ 11 ] = { h.n_vocab, h.n_audio_ctx, h.n_audio_state, h.n_audio_head, h.n_audio_layer,
		h.n_text_ctx, h.n_text_state, h.n_text_head, h.n_text_layer, h.n_mels, h.f16 };
	memcpy( out11, v, sizeof( v ) );
}
// special token ids: eot, sot, prev, solm, not, beg, translate, transcribe
void ora_special_tokens( whisper_context* c, int32_t* out8 )
{
	const auto& v = c->vocab;
	const int32_t t[ 8 ] = { v.token_eot, v.token_sot, v.token_prev, v.token_solm, v.token_not, v.token_beg,
		whisper_vocab::token_translate, whisper_vocab::token_transcribe };
	memcpy( out8, t, sizeof( t ) );
}

int ora_pcm_to_mel( whisper_context* c, const float* pcm, int n, int threads ) { return whisper_pcm_to_mel( c, pcm, n, threads ); }
int ora_set_mel( whisper_context* c, const float* mel, int n_len, int n_mel ) { return whisper_set_mel( c, mel, n_len, n_mel ); }
int ora_mel_len( whisper_context* c ) { return c->mel.n_len; }
void ora_get_mel( whisper_context* c, float* dst ) { memcpy( dst, c->mel.data.data(), c->mel.data.size() * sizeof( float ) ); }

int ora_encode( whisper_context* c, int offset, int threads ) { return whisper_encode( c, offset, threads ); }
int ora_decode( whisper_context* c, const int32_t* tokens, int n, int n_past, int threads ) { return whisper_decode( c, tokens, n, n_past, threads ); }

// logits / probs of the last decode call: N*n_vocab floats (whisper.cpp:1855-1859)
int ora_logits_size( whisper_context* c ) { return (int)c->logits.size(); }
void ora_get_logits( whisper_context* c, float* dst ) { memcpy( dst, c->logits.data(), c->logits.size() * sizeof( float ) ); }
void ora_get_probs( whisper_context* c, float* dst ) { memcpy( dst, c->probs.data(), c->probs.size() * sizeof( float ) ); }

// out5 = { id, tid, p, pt, ptsum } as doubles-in-float slots
static void packToken( const whisper_token_data& t, int32_t* ids2, float* f3 )
{
	ids2[ 0 ] = t.id; ids2[ 1 ] = t.tid; f3[ 0 ] = t.p; f3[ 1 ] = t.pt; f3[ 2 ] = t.ptsum;
}
void ora_sample_best( whisper_context* c, int32_t* ids2, float* f3 ) { packToken( whisper_sample_best( c ), ids2, f3 ); }
// GPT-2 pre-split + greedy longest match (whisper.cpp:2192-2245, 2378-2391); returns the token count or -1
int ora_tokenize( whisper_context* c, const char* text, int32_t* tokens, int cap ) { return whisper_tokenize( c, text, tokens, cap ); }
void ora_sample_timestamp( whisper_context* c, int is_initial, int32_t* ids2, float* f3 ) { packToken( whisper_sample_timestamp( c, is_initial != 0 ), ids2, f3 ); }

// sampler alone on a caller-supplied probability row: writes it where whisper_sample_best / _timestamp read (the last n_vocab
// entries of ctx->probs, whisper.cpp:2358-2376) and runs the reference's own rules (whisper.cpp:1875-1964)
void ora_sample_from_probs( whisper_context* c, const float* probs, int force_timestamp, int is_initial, int32_t* ids2, float* f3 )
{
	const int n = c->vocab.n_vocab;
	if( (int)c->probs.size() < n ) c->probs.resize( n );
	memcpy( c->probs.data() + ( c->probs.size() - n ), probs, sizeof( float ) * n );
	packToken( force_timestamp ? whisper_sample_timestamp( c, is_initial != 0 ) : whisper_sample_best( c ), ids2, f3 );
}
// language auto-detection (whisper.cpp:2428-2495): returns the language id, fills probs[whisper_lang_max_id()+1] when not null
int ora_lang_auto_detect( whisper_context* c, int offset_ms, int threads, float* lang_probs ) { return whisper_lang_auto_detect( c, offset_ms, threads, lang_probs ); }
int ora_lang_max_id() { return whisper_lang_max_id(); }

// f16 cross-attention memories written by whisper_encode (whisper.cpp:1479-1485), as f32: [n_text_layer][n_ctx][n_state]
int64_t ora_cross_kv_elements( whisper_context* c ) { return ggml_nelements( c->model.memory_cross_k ); }
void ora_get_cross_kv( whisper_context* c, float* k, float* v )
{
	const int64_t n = ggml_nelements( c->model.memory_cross_k );
	const ggml_fp16_t* pk = (const ggml_fp16_t*)c->model.memory_cross_k->data;
	const ggml_fp16_t* pv = (const ggml_fp16_t*)c->model.memory_cross_v->data;
	for( int64_t i = 0; i < n; i++ ) { k[ i ] = ggml_fp16_to_fp32( pk[ i ] ); v[ i ] = ggml_fp16_to_fp32( pv[ i ] ); }
}
int64_t ora_self_kv_elements( whisper_context* c ) { return ggml_nelements( c->model.memory_k ); }
void ora_get_self_kv( whisper_context* c, float* k, float* v )
{
	const int64_t n = ggml_nelements( c->model.memory_k );
	const ggml_fp16_t* pk = (const ggml_fp16_t*)c->model.memory_k->data;
	const ggml_fp16_t* pv = (const ggml_fp16_t*)c->model.memory_v->data;
	for( int64_t i = 0; i < n; i++ ) { k[ i ] = ggml_fp16_to_fp32( pk[ i ] ); v[ i ] = ggml_fp16_to_fp32( pv[ i ] ); }
}

// ---- decision-margin log: one entry per whisper_decode call while enabled (also inside whisper_full) ----
void ora_gap_log_enable( whisper_context* c, int on ) { g_gapLogOn = on != 0; g_gapVocab = c ? c->vocab.n_vocab : 0; g_gapLog.clear(); }
int ora_gap_log_count() { return (int)g_gapLog.size(); }
void ora_gap_log_get( float* dst ) { memcpy( dst, g_gapLog.data(), g_gapLog.size() * sizeof( float ) ); }

// ---- trace access ----
void ora_trace_enable( int on ) { g_traceOn = on != 0; g_trace.clear(); g_delayed.clear(); }
int ora_trace_count() { return (int)g_trace.size(); }
const char* ora_trace_name( int i ) { return g_trace[ i ].name.c_str(); }
void ora_trace_shape( int i, int64_t* ne4 ) { memcpy( ne4, g_trace[ i ].ne, sizeof( int64_t ) * 4 ); }
int64_t ora_trace_size( int i ) { return (int64_t)g_trace[ i ].data.size(); }
void ora_trace_data( int i, float* dst ) { memcpy( dst, g_trace[ i ].data.data(), g_trace[ i ].data.size() * sizeof( float ) ); }

// ---- whisper_full (greedy) ----
// flags: bit0 translate, bit1 no_context, bit2 single_segment, bit3 print_special
int ora_full( whisper_context* c, const float* pcm, int n, int threads, int flags, const char* language, int max_tokens, int audio_ctx )
{
	whisper_full_params p = whisper_full_default_params( WHISPER_SAMPLING_GREEDY );
	p.n_threads = threads;
	p.translate = ( flags & 1 ) != 0;
	p.no_context = ( flags & 2 ) != 0;
	p.single_segment = ( flags & 4 ) != 0;
	p.print_special = ( flags & 8 ) != 0;
	p.print_progress = false;
	p.print_realtime = false;
	p.language = language;
	p.max_tokens = max_tokens;
	p.audio_ctx = audio_ctx;
	return whisper_full( c, p, pcm, n );
}
int ora_full_ex( whisper_context* c, const float* pcm, int n, int threads, int flags, const char* language, int max_tokens, int audio_ctx, int offset_ms, int duration_ms )
{
	whisper_full_params p = whisper_full_default_params( WHISPER_SAMPLING_GREEDY );
	p.n_threads = threads;
	p.translate = ( flags & 1 ) != 0;
	p.no_context = ( flags & 2 ) != 0;
	p.single_segment = ( flags & 4 ) != 0;
	p.print_special = ( flags & 8 ) != 0;
	p.print_progress = false;
	p.print_realtime = false;
	p.language = language;
	p.max_tokens = max_tokens;
	p.audio_ctx = audio_ctx;
	p.offset_ms = offset_ms;
	p.duration_ms = duration_ms;
	return whisper_full( c, p, pcm, n );
}
// + token-level timestamps (flags bit 8 = eFullParamsFlags::TokenTimestamps) and max_len wrapping (whisper.cpp:3063-3070)
int ora_full_ex2( whisper_context* c, const float* pcm, int n, int threads, int flags, const char* language, int max_tokens, int offset_ms, int duration_ms, int max_len )
{
	whisper_full_params p = whisper_full_default_params( WHISPER_SAMPLING_GREEDY );
	p.n_threads = threads;
	p.translate = ( flags & 1 ) != 0;
	p.no_context = ( flags & 2 ) != 0;
	p.single_segment = ( flags & 4 ) != 0;
	p.print_special = ( flags & 8 ) != 0;
	p.token_timestamps = ( flags & 0x100 ) != 0;
	p.print_progress = false;
	p.print_realtime = false;
	p.language = language;
	p.max_tokens = max_tokens;
	p.max_len = max_len;
	p.offset_ms = offset_ms;
	p.duration_ms = duration_ms;
	return whisper_full( c, p, pcm, n );
}
int64_t ora_full_token_t0( whisper_context* c, int i, int j ) { return whisper_full_get_token_data( c, i, j ).t0; }
int64_t ora_full_token_t1( whisper_context* c, int i, int j ) { return whisper_full_get_token_data( c, i, j ).t1; }
float ora_full_token_vlen( whisper_context* c, int i, int j ) { return whisper_full_get_token_data( c, i, j ).vlen; }
int ora_full_n_segments( whisper_context* c ) { return whisper_full_n_segments( c ); }
int64_t ora_full_segment_t0( whisper_context* c, int i ) { return whisper_full_get_segment_t0( c, i ); }
int64_t ora_full_segment_t1( whisper_context* c, int i ) { return whisper_full_get_segment_t1( c, i ); }
const char* ora_full_segment_text( whisper_context* c, int i ) { return whisper_full_get_segment_text( c, i ); }
int ora_full_n_tokens( whisper_context* c, int i ) { return whisper_full_n_tokens( c, i ); }
int ora_full_token_id( whisper_context* c, int i, int j ) { return whisper_full_get_token_id( c, i, j ); }
float ora_full_token_p( whisper_context* c, int i, int j ) { return whisper_full_get_token_p( c, i, j ); }
void ora_clear_prompt_past( whisper_context* c ) { c->prompt_past.clear(); }

// ---- timed fixed-length chunk: the CPU baseline of bench.py ----
// pcm_to_mel + encode + [prompt decode + (n_decode-1) single-token decodes], greedy via whisper_sample_best,
// exactly the stage list BASELINE.md §2 names.  Returns wall seconds; stage_ms[3] = { mel, encode, decode+sample }.
double ora_bench_chunk( whisper_context* c, const float* pcm, int n, int threads, const int32_t* prompt, int n_prompt, int n_decode, int32_t* tokens_out, double* stage_ms )
{
	using clk = std::chrono::steady_clock;
	const auto t0 = clk::now();
	if( whisper_pcm_to_mel( c, pcm, n, threads ) != 0 ) return -1;
	const auto t1 = clk::now();
	if( whisper_encode( c, 0, threads ) != 0 ) return -2;
	const auto t2 = clk::now();
	std::vector<whisper_token> cur( prompt, prompt + n_prompt );
	int n_past = 0;
	for( int i = 0; i < n_decode; i++ )
	{
		if( whisper_decode( c, cur.data(), (int)cur.size(), n_past, threads ) != 0 ) return -3;
		n_past += (int)cur.size();
		const whisper_token_data t = ( i == 0 ) ? whisper_sample_timestamp( c, true ) : whisper_sample_best( c );
		if( tokens_out ) tokens_out[ i ] = t.id;
		cur.assign( 1, t.id );
	}
	const auto t3 = clk::now();
	auto ms = []( clk::time_point a, clk::time_point b ) { return std::chrono::duration<double, std::milli>( b - a ).count(); };
	if( stage_ms ) { stage_ms[ 0 ] = ms( t0, t1 ); stage_ms[ 1 ] = ms( t1, t2 ); stage_ms[ 2 ] = ms( t2, t3 ); }
	return ms( t0, t3 ) / 1000.0;
}

} // extern "C"
