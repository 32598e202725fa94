// Multi-GPU dispatcher, in-process: one engine + context per device and a host work queue over independent 30 s chunks.
//
// The path shards by chunk with no exchange step, so "multi-GPU" is: (1) get the weights onto every device —
// the ggml file is read ONCE, its image uploaded to the first device and copied from there to every other device over NVLink
// (cudaMemcpyPeerAsync), each engine then builds from its device-resident image; (2) hand batches of chunks to whichever replica is
// free — a queue, not a static split, so uneven clip lengths balance themselves; (3) if a replica fails, retire it and give its batch
// back to the queue.  The reference's analogues: iModel::clone shares one model between contexts
// (Whisper/Whisper/ModelImpl.cpp:40-60), whisper_full_parallel splits one clip over several states (Whisper/source/whisper.cpp:3127-3268).
// bench.py's torchrun / NCCL launch does the same thing across processes; this is the plain-C++ form of it behind the C ABI.
#include "../../include/whisper_b200.h"
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cuda_runtime.h>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace wsp
{
	int fail( int status, const std::string& what );
}

struct wsp_replicas
{
	struct Replica
	{
		int device = 0;
		wsp_engine* engine = nullptr;
		wsp_context* context = nullptr;
		std::atomic<int> failNext{ 0 };
		bool retired = false;
		float loadMs = 0;
	};
	std::vector<Replica*> reps;
	int maxBatch = 0;
	std::mutex runLock;   // one run at a time per replica set
	~wsp_replicas()
	{
		for( Replica* r : reps )
		{
			if( r->context ) wsp_context_destroy( r->context );
			if( r->engine ) wsp_engine_destroy( r->engine );
			delete r;
		}
	}
};

extern "C" {

wsp_status wsp_replicas_create( const wsp_model* m, const int32_t* devices, int32_t n_devices, int32_t max_batch, wsp_replicas** out )
{
	if( !m || !devices || !out ) return wsp::fail( WSP_E_POINTER, "model/devices/out" );
	if( n_devices < 1 || n_devices > 64 || max_batch < 1 ) return wsp::fail( WSP_E_INVALIDARG, "n_devices / max_batch" );
	uint64_t imageSize = 0;
	const void* image = wsp_model_file_data( m, &imageSize );
	if( !image || !imageSize ) return wsp::fail( WSP_E_INVALIDARG, "the model has no host file image (it was built from a meta blob)" );
	std::unique_ptr<wsp_replicas> set( new wsp_replicas() );
	set->maxBatch = max_batch;
	using clk = std::chrono::steady_clock;
	// first device: from the host image
	void* rootImage = nullptr;
	const int root = devices[ 0 ];
	for( int i = 0; i < n_devices; i++ )
	{
		wsp_replicas::Replica* r = new wsp_replicas::Replica();
		set->reps.push_back( r );
		r->device = devices[ i ];
		const auto t0 = clk::now();
		wsp_status st;
		if( n_devices == 1 ) st = wsp_engine_create( m, r->device, &r->engine );
		else
		{
			// device-resident copy of the file image: uploaded once (replica 0), then peer-to-peer from the root device
			if( cudaSetDevice( r->device ) != cudaSuccess ) return wsp::fail( WSP_E_CUDA, "cudaSetDevice" );
			void* img = nullptr;
			if( cudaMalloc( &img, imageSize ) != cudaSuccess ) return wsp::fail( WSP_E_OUTOFMEMORY, "device copy of the model file image" );
			cudaError_t ce;
			if( i == 0 ) { ce = cudaMemcpy( img, image, imageSize, cudaMemcpyHostToDevice ); rootImage = img; }
			else if( r->device == root ) ce = cudaMemcpy( img, rootImage, imageSize, cudaMemcpyDeviceToDevice );
			else
			{
				int can = 0;
				cudaDeviceCanAccessPeer( &can, r->device, root );
				if( can ) { cudaError_t pe = cudaDeviceEnablePeerAccess( root, 0 ); if( pe == cudaErrorPeerAccessAlreadyEnabled ) cudaGetLastError(); }
				ce = cudaMemcpyPeer( img, r->device, rootImage, root, imageSize );   // NVLink when peer access is on, staged through the host otherwise
			}
			if( ce != cudaSuccess ) { cudaFree( img ); return wsp::fail( WSP_E_CUDA, std::string( "model image copy: " ) + cudaGetErrorString( ce ) ); }
			st = wsp_engine_create_from_image( m, r->device, img, imageSize, &r->engine );
			if( i != 0 ) { cudaSetDevice( r->device ); cudaFree( img ); }
		}
		if( st < 0 ) { if( rootImage ) { cudaSetDevice( root ); cudaFree( rootImage ); } return st; }
		st = wsp_context_create( r->engine, max_batch, &r->context );
		if( st < 0 ) { if( rootImage ) { cudaSetDevice( root ); cudaFree( rootImage ); } return st; }
		r->loadMs = std::chrono::duration<float, std::milli>( clk::now() - t0 ).count();
	}
	if( rootImage ) { cudaSetDevice( root ); cudaFree( rootImage ); }
	*out = set.release();
	return WSP_OK;
}

int32_t wsp_replicas_count( const wsp_replicas* r ) { return r ? (int32_t)r->reps.size() : 0; }

wsp_status wsp_replicas_debug_fail_next( wsp_replicas* r, int32_t replica )
{
	if( !r || replica < 0 || replica >= (int)r->reps.size() ) return wsp::fail( WSP_E_INVALIDARG, "replica" );
	r->reps[ replica ]->failNext.store( 1 );
	return WSP_OK;
}

wsp_status wsp_replicas_run_chunks( wsp_replicas* set, const float* const* pcm, const int32_t* n_samples, int32_t n_chunks, const int32_t* prompt, int32_t n_prompt,
	int32_t n_decode, int32_t* tokens_out, wsp_replica_stats* stats )
{
	if( !set || !pcm || !n_samples || !prompt || !tokens_out ) return wsp::fail( WSP_E_POINTER, "replicas/pcm/prompt/tokens" );
	if( n_chunks < 1 || n_decode < 1 ) return wsp::fail( WSP_E_INVALIDARG, "n_chunks / n_decode" );
	std::lock_guard<std::mutex> runGuard( set->runLock );
	struct Batch { int first, count; };
	std::deque<Batch> queue;
	for( int i = 0; i < n_chunks; i += set->maxBatch ) queue.push_back( { i, std::min( set->maxBatch, n_chunks - i ) } );
	std::mutex qLock;
	std::atomic<int> remaining{ (int)queue.size() };
	std::string firstError;
	const size_t nRep = set->reps.size();
	std::vector<wsp_replica_stats> st( nRep );
	auto worker = [ & ]( size_t ri ) {
		wsp_replicas::Replica& r = *set->reps[ ri ];
		wsp_replica_stats& s = st[ ri ];
		s = wsp_replica_stats{};
		s.device = r.device;
		s.load_ms = r.loadMs;
		if( r.retired ) { s.failed = 1; return; }
		using clk = std::chrono::steady_clock;
		while( remaining.load() > 0 )
		{
			Batch b;
			{
				std::lock_guard<std::mutex> lk( qLock );
				if( queue.empty() )
				{
					// nothing to take right now, but a peer may still fail and give its batch back
					b.count = 0;
				}
				else { b = queue.front(); queue.pop_front(); }
			}
			if( b.count == 0 ) { std::this_thread::sleep_for( std::chrono::microseconds( 200 ) ); continue; }
			const auto t0 = clk::now();
			wsp_status rc = r.failNext.exchange( 0 ) ? (wsp_status)WSP_E_CUDA
				: wsp_run_chunks( r.context, pcm + b.first, n_samples + b.first, b.count, prompt, n_prompt, n_decode, tokens_out + (size_t)b.first * n_decode, nullptr );
			s.busy_ms += std::chrono::duration<float, std::milli>( clk::now() - t0 ).count();
			if( rc < 0 )
			{
				// retire this replica; its batch goes back to the queue for the others
				std::lock_guard<std::mutex> lk( qLock );
				if( firstError.empty() ) firstError = wsp_last_error();
				queue.push_front( b );
				r.retired = true;
				s.failed = 1;
				return;
			}
			s.batches_done++;
			s.chunks_done += b.count;
			remaining.fetch_sub( 1 );
		}
	};
	std::vector<std::thread> threads;
	for( size_t i = 1; i < nRep; i++ ) threads.emplace_back( worker, i );
	worker( 0 );
	// if replica 0 retired early the others keep draining the queue; wait for them, then see what is left
	for( auto& t : threads ) t.join();
	if( stats ) for( size_t i = 0; i < nRep; i++ ) stats[ i ] = st[ i ];
	if( remaining.load() > 0 )
		return wsp::fail( WSP_E_CUDA, "every replica failed; first error: " + firstError );
	return WSP_OK;
}

void wsp_replicas_destroy( wsp_replicas* r ) { delete r; }

} // extern "C"
