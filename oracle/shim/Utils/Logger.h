/* TEST INFRASTRUCTURE — build shim for the parity oracle (oracle/_ref).
 *
 * The reference's own CPU implementation (Whisper/source/ggml.c + whisper.cpp) is compiled
 * UNMODIFIED from /root/reference.  whisper.cpp does `#include "Utils/Logger.h"`
 * (Whisper/source/whisper.cpp:18) and calls Tracing::* at its named trace points
 * (whisper.cpp:1121-1869); the real headers (Whisper/Utils/Logger.h, Whisper/Utils/Trace/tracing.h)
 * pull in <windows.h>/D3D types.  This file stands in for both: it only DECLARES the four printf-style
 * loggers and a Tracing namespace whose functions record named tensors into a table that
 * oracle/ref_wrap.cpp exposes through ora_trace_*().  No reference logic is restated here. */
#pragma once
#include <stdio.h>
#include <stdarg.h>
#ifdef __cplusplus
typedef char8_t ora_fmt_t;   /* whisper.cpp passes u8"..." literals: char8_t under -std=c++20 */
extern "C" {
#else
typedef char ora_fmt_t;      /* ggml.c is C: u8"..." is plain char */
#endif
void logError( const ora_fmt_t* fmt, ... );
void logWarning( const ora_fmt_t* fmt, ... );
void logInfo( const ora_fmt_t* fmt, ... );
void logDebug( const ora_fmt_t* fmt, ... );
#ifdef __cplusplus
}
#include <vector>
struct ggml_tensor;
namespace Tracing
{
	/* Same call signatures as Whisper/Utils/Trace/tracing.h:56-70 (the no-op branch). */
	struct ItemName
	{
		char text[ 96 ];
		ItemName( const char* s ) { snprintf( text, sizeof( text ), "%s", s ); }
		ItemName( const char* f, int i ) { snprintf( text, sizeof( text ), f, i ); }
	};
	void tensor( const ItemName& name, const ggml_tensor* t );
	void delayTensor( const ItemName& name, const ggml_tensor* t );
	void writeDelayedTensors();
	void vector( const ItemName& name, const std::vector<float>& v );
}
#endif
