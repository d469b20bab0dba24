// TEST INFRASTRUCTURE ONLY: pansvr::GlibcRandom against the host libc's rand() and random_r().
#include "../../pansvr_b200/csrc/aln/refrand.hpp"
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <initializer_list>
int main()
{
	pansvr::GlibcRandom g;
	int bad = 0;
	for (int i = 0; i < 200000; ++i) if (rand() != g.next()) ++bad;
	struct random_data rd;
	char st[128];
	for (unsigned seed : {1u, 12345u, 1804289383u, 0u}) {
		memset(&rd, 0, sizeof rd); memset(st, 0, sizeof st);
		initstate_r(seed, st, 128, &rd);
		pansvr::GlibcRandom h(seed);
		for (int i = 0; i < 100000; ++i) { int32_t x; random_r(&rd, &x); if (x != h.next()) ++bad; }
	}
	if (bad) fprintf(stderr, "refrand: %d mismatches\n", bad);
	return bad != 0;
}
