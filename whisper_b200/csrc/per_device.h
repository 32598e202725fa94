// cudaFuncSetAttribute (dynamic shared memory opt-in) applies to the CURRENT device only, and several context threads may launch
// at once: every "has the attribute been raised yet" memo is kept per device under a mutex.
#pragma once
#include <cuda_runtime.h>
#include <mutex>
#include <stddef.h>

namespace kern
{
	class PerDeviceMax
	{
		std::mutex m;
		size_t v[ 64 ] = {};

	public:
		// calls set( need ) when `need` exceeds what was set so far on the current device
		template<class F>
		cudaError_t raise( size_t need, F set )
		{
			int dev = 0;
			cudaError_t e = cudaGetDevice( &dev );
			if( e != cudaSuccess ) return e;
			if( dev < 0 || dev >= 64 ) return cudaErrorInvalidDevice;
			std::lock_guard<std::mutex> lk( m );
			if( need <= v[ dev ] ) return cudaSuccess;
			e = set( need );
			if( e == cudaSuccess ) v[ dev ] = need;
			return e;
		}
	};
}
