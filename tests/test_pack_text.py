"""The sixteen-at-a-time text packer (tagdust_b200/csrc/tdg_pack.h) against its own table-driven path: every byte value,
every length around the block sizes.  Host-only: compiled with g++ here, no GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r"""
#include "tdg_pack.h"
#include <cstdio>
#include <cstdlib>
#include <vector>
int main()
{
	uint8_t table[256];
	for (int c = 0; c < 256; c++) table[c] = 4;
	table['A'] = table['a'] = 0; table['C'] = table['c'] = 1; table['G'] = table['g'] = 2;
	table['T'] = table['t'] = table['U'] = table['u'] = 3; table['.'] = 5;
	if (!tdg::is_nuc_code_table(table)) { printf("table not recognised\n"); return 1; }
	table['x'] = 0;
	if (tdg::is_nuc_code_table(table)) { printf("altered table recognised\n"); return 1; }
	table['x'] = 4;
	unsigned long long seed = 12345;
	auto rnd = [&] { seed = seed * 6364136223846793005ULL + 1442695040888963407ULL; return (unsigned)(seed >> 33); };
	const char* common = "ACGTNacgtnUu.";
	long checked = 0;
	for (int len = 0; len <= 210; len++)
		for (int rep = 0; rep < 60; rep++) {
			std::vector<uint8_t> s((size_t)len + 32);
			for (auto& c : s) c = (rep % 3 == 0) ? (uint8_t)(32 + rnd() % 224) : (uint8_t)common[rnd() % 13];
			const int max_len = len + (int)(rnd() % 20);
			const int words = (max_len + 1 + 7) / 8;
			std::vector<uint32_t> a((size_t)words * 32, 0xDEADBEEFu), b((size_t)words * 32, 0xDEADBEEFu);
			const int lane = (int)(rnd() % 32);
			tdg::pack_text_words(a.data() + lane, words, s.data(), len, table, true);
			tdg::pack_text_words(b.data() + lane, words, s.data(), len, table, false);
			if (a != b) { printf("mismatch at len %d rep %d\n", len, rep); return 1; }
			for (int w = 0; w < words; w++)
				for (int k = 0; k < 8; k++) {
					const int pos = w * 8 + k;
					const unsigned want = pos < len ? table[s[(size_t)pos]] : 0u;
					if (((a[(size_t)w * 32 + lane] >> (4 * k)) & 0xF) != want) { printf("wrong code at len %d pos %d\n", len, pos); return 1; }
				}
			checked++;
		}
	printf("ok %ld\n", checked);
	return 0;
}
"""


def test_pack_text_words_simd_equals_table(tmp_path):
    src = tmp_path / "pack_test.cpp"
    src.write_text(SRC)
    exe = tmp_path / "pack_test"
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "tagdust_b200", "csrc"), str(src), "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout + r.stderr
