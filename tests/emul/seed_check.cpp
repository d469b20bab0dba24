// TEST INFRASTRUCTURE ONLY: the word-wise MEM extension of seed_core.cuh against the reference's base-by-base loops
// (deBGA_index.cpp:118-131) on random packed sequences.  Exit code 0 = identical.
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../../pansvr_b200/csrc/aln/seed_core.cuh"

using namespace pansvr;

static uint64_t rng_state = 88172645463325252ull;
static uint64_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }

int main()
{
	const size_t NB = 4096;                                     // bases per sequence
	std::vector<uint64_t> a(NB / 32 + 1, 0), b(NB / 32 + 1, 0);
	long checked = 0;
	for (int round = 0; round < 400; ++round) {
		for (size_t i = 0; i < NB / 32; ++i) { a[i] = rnd(); b[i] = rnd(); }
		// plant shared runs of assorted lengths so that long matches, word-boundary crossings and caps are all exercised
		for (int k = 0; k < 40; ++k) {
			const size_t len = 1 + rnd() % 200, pa = rnd() % (NB - len), pb = rnd() % (NB - len);
			for (size_t i = 0; i < len; ++i) {
				const uint64_t base = (a[(pa + i) >> 5] >> ((31 - ((pa + i) & 31)) << 1)) & 3;
				uint64_t &w = b[(pb + i) >> 5];
				const int sh = (31 - ((pb + i) & 31)) << 1;
				w = (w & ~(3ull << sh)) | (base << sh);
			}
			for (int t = 0; t < 50; ++t) {
				const size_t d = rnd() % len;
				const uint64_t ap = pa + d, bp = pb + d;
				// right
				uint32_t cap = (uint32_t)(rnd() % 260);
				if (ap + cap > NB) cap = (uint32_t)(NB - ap);
				if (bp + cap > NB) cap = (uint32_t)(NB - bp);
				uint32_t naive = 0;
				while (naive < cap && base_at(a.data(), ap + naive) == base_at(b.data(), bp + naive)) ++naive;
				if (match_right(a.data(), ap, b.data(), bp, cap) != naive) { fprintf(stderr, "match_right differs\n"); return 1; }
				// left
				cap = (uint32_t)(rnd() % 260);
				if (cap > ap) cap = (uint32_t)ap;
				if (cap > bp) cap = (uint32_t)bp;
				naive = 0;
				while (naive < cap && base_at(a.data(), ap - 1 - naive) == base_at(b.data(), bp - 1 - naive)) ++naive;
				if (match_left(a.data(), ap, b.data(), bp, cap) != naive) { fprintf(stderr, "match_left differs\n"); return 1; }
				checked += 2;
			}
		}
	}
	printf("%ld extensions identical\n", checked);
	return 0;
}
