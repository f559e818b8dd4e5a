// Packing of one read from the characters of a FASTQ / FASTA text into the device layout: 4-bit codes, eight per 32-bit
// word, word w of the 32 reads of a tile next to each other (stride 32 words).  Host-only; shared by tdg_host.cu and the
// host-pipeline measurement harness (scripts/micro/host_pipeline.cpp).
#pragma once
#include <cstdint>
#include <cstring>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace tdg {

// true where code_of[] is the reference's nuc_code table (nuc_code.c:46-74): A/a 0, C/c 1, G/g 2, T/t/U/u 3, '.' 5,
// everything else 4 -- the mapping the sixteen-at-a-time path below computes without the table
inline bool is_nuc_code_table(const uint8_t* code_of)
{
	for (int c = 0; c < 256; c++) {
		uint8_t want = 4;
		switch (c) {
			case 'A': case 'a': want = 0; break;
			case 'C': case 'c': want = 1; break;
			case 'G': case 'g': want = 2; break;
			case 'T': case 't': case 'U': case 'u': want = 3; break;
			case '.': want = 5; break;
			default: break;
		}
		if (code_of[c] != want) return false;
	}
	return true;
}

// base: word 0 of the read inside its tile (stride 32 words); s[0..len) the characters.  The code behind the last base is 0.
inline void pack_text_words(uint32_t* base, int words, const uint8_t* s, int len, const uint8_t* code_of, bool nuc_table)
{
	int pos = 0, w = 0;
#if defined(__SSE2__)
	if (nuc_table) {
		const __m128i kDF = _mm_set1_epi8((char)0xDF), k3 = _mm_set1_epi8(3), k4 = _mm_set1_epi8(4), k1 = _mm_set1_epi8(1);
		const __m128i cA = _mm_set1_epi8('A'), cC = _mm_set1_epi8('C'), cG = _mm_set1_epi8('G'), cT = _mm_set1_epi8('T'), cU = _mm_set1_epi8('U');
		const __m128i cDot = _mm_set1_epi8('.'), kLo = _mm_set1_epi16(0x00FF);
		for (; pos + 16 <= len && w + 2 <= words; pos += 16, w += 2) {
			const __m128i x = _mm_loadu_si128((const __m128i*)(s + pos));
			const __m128i f = _mm_and_si128(x, kDF);   // letters in upper case; nothing else becomes a letter
			const __m128i valid = _mm_or_si128(_mm_or_si128(_mm_or_si128(_mm_cmpeq_epi8(f, cA), _mm_cmpeq_epi8(f, cC)), _mm_or_si128(_mm_cmpeq_epi8(f, cG), _mm_cmpeq_epi8(f, cT))),
			                                   _mm_cmpeq_epi8(f, cU));
			// A 0x41, C 0x43, G 0x47, T 0x54, U 0x55: bits 1-2 xor bits 2-3 of the character = 0, 1, 2, 3, 3
			const __m128i c2 = _mm_and_si128(_mm_xor_si128(_mm_srli_epi16(x, 1), _mm_srli_epi16(x, 2)), k3);
			__m128i code = _mm_or_si128(_mm_and_si128(c2, valid), _mm_andnot_si128(valid, k4));
			code = _mm_or_si128(code, _mm_and_si128(_mm_cmpeq_epi8(x, cDot), k1));   // '.': 4 | 1
			// nibbles: byte 2j | byte 2j+1 << 4
			const __m128i w16 = _mm_or_si128(_mm_and_si128(code, kLo), _mm_slli_epi16(_mm_srli_epi16(code, 8), 4));
			const uint64_t two = (uint64_t)_mm_cvtsi128_si64(_mm_packus_epi16(w16, w16));
			base[(size_t)w * 32] = (uint32_t)two;
			base[(size_t)(w + 1) * 32] = (uint32_t)(two >> 32);
		}
	}
#else
	(void)nuc_table;
#endif
	for (; w < words; w++) {
		uint32_t v = 0;
		if (pos + 8 <= len) {
			for (int k = 0; k < 8; k++) v |= (uint32_t)code_of[s[pos + k]] << (4 * k);
			pos += 8;
		} else {
			for (int k = 0; k < 8 && pos < len; k++, pos++) v |= (uint32_t)(code_of[s[pos]] & 0xF) << (4 * k);
			pos = len + 8;   // the terminator (code 0) and everything behind it are zero bits
		}
		base[(size_t)w * 32] = v;
	}
}

// Where the reads of a text chunk go: the pinned staging arrays of a batch (tdg::batch_text_target in tdg_host.cu).
// Reads first .. first + n - 1 of the batch may be packed by any number of threads at once, each read by one of them.
struct TextTarget {
	uint32_t* seq = nullptr;   // tiles of 32 reads x words
	int32_t*  len = nullptr;
	int words = 0, max_len = 0, first = 0;
};

inline void pack_text_read(const TextTarget& t, int i, const uint8_t* s, int len, const uint8_t* code_of, bool nuc_table)
{
	const int r = t.first + i;
	pack_text_words(t.seq + ((size_t)(r >> 5) * t.words) * 32 + (r & 31), t.words, s, len, code_of, nuc_table);
	t.len[r] = len;
}

}  // namespace tdg
