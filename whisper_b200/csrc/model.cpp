#include "model.h"
#include "../../include/whisper_b200.h"
#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace wsp
{
	ModelFile::~ModelFile()
	{
		if( image && fd >= 0 ) munmap( const_cast<uint8_t*>( image ), imageSize );
		if( fd >= 0 ) close( fd );
	}
	const TensorInfo* ModelFile::find( const std::string& name ) const
	{
		auto it = index.find( name );
		return it == index.end() ? nullptr : &tensors[ it->second ];
	}

	namespace
	{
		struct Reader
		{
			const uint8_t* p;
			uint64_t size, pos = 0;
			bool ok = true;
			template<class T> T get()
			{
				T v{};
				if( sizeof( T ) > size - pos ) { ok = false; return v; }   // (pos <= size always: overflow-safe form)
				memcpy( &v, p + pos, sizeof( T ) );
				pos += sizeof( T );
				return v;
			}
			bool skip( uint64_t n ) { if( n > size - pos ) { ok = false; return false; } pos += n; return true; }
		};

		void finishVocab( Vocab& v, int32_t nWords )
		{
			// whisper.cpp:575-607: special ids shift for multilingual models; missing entries get synthetic names
			if( v.multilingual() )
			{
				v.token_eot++; v.token_sot++; v.token_prev++; v.token_solm++; v.token_not++; v.token_beg++;
			}
			if( (int32_t)v.id_to_token.size() < v.n_vocab )
				v.id_to_token.resize( v.n_vocab );
			for( int32_t i = nWords; i < v.n_vocab; i++ )
			{
				std::string w;
				if( i > v.token_beg ) w = "[_TT_" + std::to_string( i - v.token_beg ) + "]";
				else if( i == v.token_eot ) w = "[_EOT_]";
				else if( i == v.token_sot ) w = "[_SOT_]";
				else if( i == v.token_prev ) w = "[_PREV_]";
				else if( i == v.token_not ) w = "[_NOT_]";
				else if( i == v.token_beg ) w = "[_BEG_]";
				else w = "[_extra_token_" + std::to_string( i ) + "]";
				v.id_to_token[ i ] = w;
				v.token_to_id[ w ] = i;
			}
		}
	}

	int parseModelImage( const uint8_t* data, uint64_t size, ModelFile& m, std::string& err )
	{
		Reader r{ data, size };
		if( r.get<uint32_t>() != 0x67676d6cu ) { err = "bad magic (not a ggml file)"; return WSP_E_FORMAT; }
		int32_t* hp = reinterpret_cast<int32_t*>( &m.hp );
		for( int i = 0; i < 11; i++ ) hp[ i ] = r.get<int32_t>();
		if( !r.ok ) { err = "truncated header"; return WSP_E_FILE; }
		const HParams& h = m.hp;
		if( h.n_audio_state != h.n_text_state || h.n_audio_state <= 0 || h.n_audio_state % 128 != 0 || h.n_audio_head * 64 != h.n_audio_state ||
			h.n_text_head * 64 != h.n_text_state || h.n_mels != 80 || h.n_vocab <= 0 || h.n_audio_ctx <= 0 || h.n_text_ctx <= 0 )
		{
			err = "unsupported hyper-parameters (need head dim 64, n_mels 80, state multiple of 128)";
			return WSP_E_FORMAT;
		}
		m.filt_n_mel = r.get<int32_t>();
		m.filt_n_fft = r.get<int32_t>();
		if( !r.ok || m.filt_n_mel != 80 || m.filt_n_fft != 201 ) { err = "unexpected mel filter bank shape"; return WSP_E_FORMAT; }
		m.filters.resize( (size_t)m.filt_n_mel * m.filt_n_fft );
		if( r.pos + m.filters.size() * 4 > size ) { err = "truncated filters"; return WSP_E_FILE; }
		memcpy( m.filters.data(), data + r.pos, m.filters.size() * 4 );
		r.pos += m.filters.size() * 4;

		const int32_t nWords = r.get<int32_t>();
		if( !r.ok || nWords < 0 || nWords > 1000000 ) { err = "bad vocabulary size"; return WSP_E_FORMAT; }
		Vocab& v = m.vocab;
		v.n_vocab = h.n_vocab;
		v.id_to_token.resize( nWords > h.n_vocab ? nWords : h.n_vocab );
		for( int32_t i = 0; i < nWords; i++ )
		{
			const uint32_t len = r.get<uint32_t>();
			if( !r.ok || r.pos + len > size ) { err = "truncated vocabulary"; return WSP_E_FILE; }
			std::string w( reinterpret_cast<const char*>( data + r.pos ), len );
			r.pos += len;
			v.id_to_token[ i ] = w;
			v.token_to_id[ w ] = i;
		}
		finishVocab( v, nWords );

		while( r.pos < size )
		{
			TensorInfo t;
			t.n_dims = r.get<int32_t>();
			const int32_t nameLen = r.get<int32_t>();
			t.ftype = r.get<int32_t>();
			if( !r.ok ) break;   // trailing bytes shorter than a header: treat as EOF like the reference (whisper.cpp:1012)
			if( t.n_dims < 1 || t.n_dims > 3 || nameLen <= 0 || nameLen > 256 ) { err = "corrupt tensor header"; return WSP_E_FORMAT; }
			for( int i = 0; i < t.n_dims; i++ ) t.ne[ i ] = r.get<int32_t>();
			if( !r.ok || (uint64_t)nameLen > size - r.pos ) { err = "truncated tensor header"; return WSP_E_FILE; }
			for( int i = 0; i < t.n_dims; i++ )
				if( t.ne[ i ] <= 0 ) { err = "corrupt tensor header: non-positive dimension"; return WSP_E_FORMAT; }
			if( t.ftype != 0 && t.ftype != 1 ) { err = "unsupported tensor type (only f32 = 0 and f16 = 1 are defined, whisper.cpp:1003-1051)"; return WSP_E_FORMAT; }
			t.name.assign( reinterpret_cast<const char*>( data + r.pos ), nameLen );
			r.pos += nameLen;
			t.nbytes = (uint64_t)t.elements() * ( t.ftype == 0 ? 4 : 2 );
			t.offset = r.pos;
			if( !r.skip( t.nbytes ) ) { err = "truncated tensor data: " + t.name; return WSP_E_FILE; }
			m.index[ t.name ] = (int)m.tensors.size();
			m.tensors.push_back( std::move( t ) );
		}
		return validateTensors( m, err );
	}

	int validateTensors( const ModelFile& m, std::string& err )
	{
		const HParams& h = m.hp;
		const int d = h.n_audio_state;
		auto need = [ & ]( const std::string& name, int n0, int n1, int n2 ) -> bool {
			const TensorInfo* t = m.find( name );
			if( !t ) { err = "missing tensor " + name; return false; }
			if( t->ne[ 0 ] != n0 || t->ne[ 1 ] != n1 || t->ne[ 2 ] != n2 ) { err = "tensor " + name + " has wrong shape"; return false; }
			return true;
		};
		bool ok = need( "encoder.positional_embedding", d, h.n_audio_ctx, 1 ) && need( "encoder.conv1.weight", 3, h.n_mels, d ) &&
			need( "encoder.conv1.bias", 1, d, 1 ) && need( "encoder.conv2.weight", 3, d, d ) && need( "encoder.conv2.bias", 1, d, 1 ) &&
			need( "encoder.ln_post.weight", d, 1, 1 ) && need( "encoder.ln_post.bias", d, 1, 1 ) &&
			need( "decoder.positional_embedding", d, h.n_text_ctx, 1 ) && need( "decoder.token_embedding.weight", d, h.n_vocab, 1 ) &&
			need( "decoder.ln.weight", d, 1, 1 ) && need( "decoder.ln.bias", d, 1, 1 );
		auto block = [ & ]( const std::string& p, bool cross ) -> bool {
			bool b = need( p + "attn_ln.weight", d, 1, 1 ) && need( p + "attn_ln.bias", d, 1, 1 ) &&
				need( p + "attn.query.weight", d, d, 1 ) && need( p + "attn.query.bias", d, 1, 1 ) && need( p + "attn.key.weight", d, d, 1 ) &&
				need( p + "attn.value.weight", d, d, 1 ) && need( p + "attn.value.bias", d, 1, 1 ) &&
				need( p + "attn.out.weight", d, d, 1 ) && need( p + "attn.out.bias", d, 1, 1 ) &&
				need( p + "mlp_ln.weight", d, 1, 1 ) && need( p + "mlp_ln.bias", d, 1, 1 ) &&
				need( p + "mlp.0.weight", d, 4 * d, 1 ) && need( p + "mlp.0.bias", 4 * d, 1, 1 ) &&
				need( p + "mlp.2.weight", 4 * d, d, 1 ) && need( p + "mlp.2.bias", d, 1, 1 );
			if( b && cross )
				b = need( p + "cross_attn_ln.weight", d, 1, 1 ) && need( p + "cross_attn_ln.bias", d, 1, 1 ) &&
					need( p + "cross_attn.query.weight", d, d, 1 ) && need( p + "cross_attn.query.bias", d, 1, 1 ) &&
					need( p + "cross_attn.key.weight", d, d, 1 ) && need( p + "cross_attn.value.weight", d, d, 1 ) &&
					need( p + "cross_attn.value.bias", d, 1, 1 ) && need( p + "cross_attn.out.weight", d, d, 1 ) && need( p + "cross_attn.out.bias", d, 1, 1 );
			return b;
		};
		for( int i = 0; ok && i < h.n_audio_layer; i++ ) ok = block( "encoder.blocks." + std::to_string( i ) + ".", false );
		for( int i = 0; ok && i < h.n_text_layer; i++ ) ok = block( "decoder.blocks." + std::to_string( i ) + ".", true );
		return ok ? WSP_OK : WSP_E_FORMAT;
	}

	int openModelFile( const char* path, ModelFile** out, std::string& err )
	{
		const int fd = open( path, O_RDONLY );
		if( fd < 0 ) { err = std::string( "cannot open " ) + path; return WSP_E_FILE; }
		struct stat st;
		if( fstat( fd, &st ) != 0 || st.st_size < 64 ) { close( fd ); err = "cannot stat / file too small"; return WSP_E_FILE; }
		void* p = mmap( nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0 );
		if( p == MAP_FAILED ) { close( fd ); err = "mmap failed"; return WSP_E_FILE; }
		ModelFile* m = new ModelFile();
		m->fd = fd;
		m->image = static_cast<const uint8_t*>( p );
		m->imageSize = (uint64_t)st.st_size;
		const int rc = parseModelImage( m->image, m->imageSize, *m, err );
		if( rc != WSP_OK ) { delete m; return rc; }
		*out = m;
		return WSP_OK;
	}

	// meta blob: "WSPM" u32 version | hparams | filters | n_vocab strings | tensor directory | image size
	namespace
	{
		template<class T> void put( std::vector<uint8_t>& d, const T& v ) { const uint8_t* p = reinterpret_cast<const uint8_t*>( &v ); d.insert( d.end(), p, p + sizeof( T ) ); }
		void putStr( std::vector<uint8_t>& d, const std::string& s ) { put<uint32_t>( d, (uint32_t)s.size() ); d.insert( d.end(), s.begin(), s.end() ); }
	}
	void serializeMeta( const ModelFile& m, std::vector<uint8_t>& d )
	{
		d.clear();
		put<uint32_t>( d, 0x4d505357u );
		put<uint32_t>( d, 1 );
		put( d, m.hp );
		put<uint32_t>( d, (uint32_t)m.filters.size() );
		const uint8_t* f = reinterpret_cast<const uint8_t*>( m.filters.data() );
		d.insert( d.end(), f, f + m.filters.size() * 4 );
		put<uint32_t>( d, (uint32_t)m.vocab.id_to_token.size() );
		for( const auto& s : m.vocab.id_to_token ) putStr( d, s );
		put<uint32_t>( d, (uint32_t)m.tensors.size() );
		for( const auto& t : m.tensors )
		{
			putStr( d, t.name );
			put( d, t.n_dims ); put( d, t.ne[ 0 ] ); put( d, t.ne[ 1 ] ); put( d, t.ne[ 2 ] ); put( d, t.ftype ); put( d, t.offset ); put( d, t.nbytes );
		}
		put<uint64_t>( d, m.imageSize );
	}
	int modelFromMeta( const uint8_t* data, uint64_t size, ModelFile** out, std::string& err )
	{
		Reader r{ data, size };
		if( r.get<uint32_t>() != 0x4d505357u || r.get<uint32_t>() != 1 ) { err = "bad meta blob"; return WSP_E_FORMAT; }
		ModelFile* m = new ModelFile();
		m->hp = r.get<HParams>();
		const uint32_t nf = r.get<uint32_t>();
		if( !r.ok || r.pos + (uint64_t)nf * 4 > size ) { delete m; err = "truncated meta"; return WSP_E_FORMAT; }
		m->filters.resize( nf );
		memcpy( m->filters.data(), data + r.pos, (size_t)nf * 4 );
		r.pos += (uint64_t)nf * 4;
		m->filt_n_mel = 80; m->filt_n_fft = 201;
		auto getStr = [ & ]() { const uint32_t n = r.get<uint32_t>(); std::string s; if( r.ok && r.pos + n <= size ) { s.assign( (const char*)data + r.pos, n ); r.pos += n; } else r.ok = false; return s; };
		const uint32_t nv = r.get<uint32_t>();
		m->vocab.n_vocab = m->hp.n_vocab;
		m->vocab.id_to_token.resize( nv );
		for( uint32_t i = 0; i < nv && r.ok; i++ ) { m->vocab.id_to_token[ i ] = getStr(); m->vocab.token_to_id[ m->vocab.id_to_token[ i ] ] = (int32_t)i; }
		if( m->vocab.multilingual() ) { m->vocab.token_eot++; m->vocab.token_sot++; m->vocab.token_prev++; m->vocab.token_solm++; m->vocab.token_not++; m->vocab.token_beg++; }
		const uint32_t nt = r.get<uint32_t>();
		for( uint32_t i = 0; i < nt && r.ok; i++ )
		{
			TensorInfo t;
			t.name = getStr();
			t.n_dims = r.get<int32_t>(); t.ne[ 0 ] = r.get<int32_t>(); t.ne[ 1 ] = r.get<int32_t>(); t.ne[ 2 ] = r.get<int32_t>();
			t.ftype = r.get<int32_t>(); t.offset = r.get<uint64_t>(); t.nbytes = r.get<uint64_t>();
			m->index[ t.name ] = (int)m->tensors.size();
			m->tensors.push_back( std::move( t ) );
		}
		m->imageSize = r.get<uint64_t>();
		if( !r.ok ) { delete m; err = "truncated meta"; return WSP_E_FORMAT; }
		// the blob is trusted for nothing: every tensor must lie inside the image it describes, with the size its shape implies
		for( const TensorInfo& t : m->tensors )
		{
			bool good = t.n_dims >= 1 && t.n_dims <= 3 && ( t.ftype == 0 || t.ftype == 1 );
			for( int i = 0; good && i < t.n_dims; i++ ) good = t.ne[ i ] > 0;
			good = good && t.nbytes == (uint64_t)t.elements() * ( t.ftype == 0 ? 4 : 2 ) && t.offset <= m->imageSize && t.nbytes <= m->imageSize - t.offset;
			if( !good ) { err = "meta blob: tensor " + t.name + " is inconsistent"; delete m; return WSP_E_FORMAT; }
		}
		const int rc = validateTensors( *m, err );
		if( rc != WSP_OK ) { delete m; return rc; }
		*out = m;
		return WSP_OK;
	}
}
