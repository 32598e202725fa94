// ggml Whisper model file: host-side parse (format: SURVEY.md Appendix A; reference readers Whisper/source/whisper.cpp:451-1072,
// Whisper/Whisper/WhisperModel.cpp:434-492, Whisper/Whisper/Vocabulary.cpp:64-143).  The file is mapped, never copied; the
// "meta" blob is the small part a peer rank needs next to a broadcast file image.
#pragma once
#include <stdint.h>
#include <map>
#include <string>
#include <vector>

namespace wsp
{
	struct HParams
	{
		int32_t n_vocab, n_audio_ctx, n_audio_state, n_audio_head, n_audio_layer;
		int32_t n_text_ctx, n_text_state, n_text_head, n_text_layer, n_mels, f16;
	};

	struct TensorInfo
	{
		std::string name;
		int32_t n_dims = 0;
		int32_t ne[ 3 ] = { 1, 1, 1 };
		int32_t ftype = 0;        // 0 = f32, else f16
		uint64_t offset = 0;      // byte offset of the data in the file image
		uint64_t nbytes = 0;
		int64_t elements() const { return (int64_t)ne[ 0 ] * ne[ 1 ] * ne[ 2 ]; }
	};

	struct Vocab
	{
		int32_t n_vocab = 0;
		std::vector<std::string> id_to_token;
		std::map<std::string, int32_t> token_to_id;
		// whisper.cpp:199-221 (+1 when multilingual, :575-583)
		int32_t token_eot = 50256, token_sot = 50257, token_prev = 50360, token_solm = 50361, token_not = 50362, token_beg = 50363;
		static constexpr int32_t token_translate = 50358, token_transcribe = 50359;
		bool multilingual() const { return n_vocab == 51865; }
	};

	struct ModelFile
	{
		HParams hp{};
		int32_t filt_n_mel = 0, filt_n_fft = 0;
		std::vector<float> filters;
		Vocab vocab;
		std::vector<TensorInfo> tensors;
		std::map<std::string, int> index;
		// mapped file image (null when built from a meta blob)
		const uint8_t* image = nullptr;
		uint64_t imageSize = 0;
		int fd = -1;

		~ModelFile();
		const TensorInfo* find( const std::string& name ) const;
	};

	// returns 0 or a negative wsp_status; err receives a description
	int openModelFile( const char* path, ModelFile** out, std::string& err );
	int parseModelImage( const uint8_t* data, uint64_t size, ModelFile& m, std::string& err );
	void serializeMeta( const ModelFile& m, std::vector<uint8_t>& dst );
	int modelFromMeta( const uint8_t* data, uint64_t size, ModelFile** out, std::string& err );
	// every tensor the network needs is present with the expected shape (whisper.cpp:1028-1067)
	int validateTensors( const ModelFile& m, std::string& err );
}
