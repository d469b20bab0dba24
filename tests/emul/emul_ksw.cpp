// emul_ksw.cpp -- TEST INFRASTRUCTURE ONLY.
// Compiles the warp program of pansvr_b200/csrc/ksw_fast.cuh for the host (PANSVR_HOST_EMUL) and
// steps it on the 32-fibre lock-step warp of warp_emul.hpp, so the kernel's logic can be checked
// against the oracle on a machine without a GPU.  Never linked into the product library.
#define PANSVR_HOST_EMUL 1
#include <vector>
#include <string.h>
#include "../../pansvr_b200/csrc/ksw_host.hpp"

template <int CPL, bool WRAP>
static void run_one(const kswfast::Params &P, int qlen, const uint8_t *q, int tlen, const uint8_t *t, int32_t *res,
                    uint32_t *cigar, int cigar_cap)
{
	const int W = 32 * CPL;
	const int nd = kswhost::n_diagonals(qlen, tlen, P.w);
	std::vector<uint8_t> tb((size_t)(nd + 1) * W + 64, 0xEE), QS(qlen + 2 + 16);
	std::vector<int32_t> Hs(W);
	WarpEmul::run([&]() { kswfast::align_task<CPL, WRAP>(P, qlen, q, tlen, t, res, cigar, cigar_cap, tb.data(), Hs.data(), QS.data()); });
}

extern "C" int emul_ksw_fast_batch(int n, const uint8_t *qseq, const int64_t *qoff, const int32_t *qlen, const uint8_t *tseq,
                                   const int64_t *toff, const int32_t *tlen, int m, const int8_t *mat, int q, int e, int q2,
                                   int e2, int w, int zdrop, int end_bonus, int flag, int32_t *res, uint32_t *cigar,
                                   int cigar_cap, int force_cpl, int force_wrap)
{
	kswhost::Plan pl = kswhost::make_plan(m, mat, q, e, q2, e2, w, zdrop, end_bonus, flag);
	if (!pl.trivial && !pl.fast_params) return -1;
	for (int i = 0; i < n; ++i) {
		int32_t *o = res + (size_t)i * kswfast::RES_WORDS;
		uint32_t *c = cigar + (size_t)i * cigar_cap;
		memset(c, 0, sizeof(uint32_t) * cigar_cap);
		if (pl.trivial || qlen[i] <= 0 || tlen[i] <= 0) {
			const int32_t z[12] = {0, 0, -1, -1, kswfast::NEG_INF, -1, kswfast::NEG_INF, -1, kswfast::NEG_INF, 0, 0, 0};
			memcpy(o, z, sizeof(z));
			continue;
		}
		int cpl = kswhost::pick_cpl(qlen[i], tlen[i], w);
		if (cpl == 0) return -2;
		if (force_cpl > cpl) cpl = force_cpl;
		const uint8_t *qq = qseq + qoff[i], *tt = tseq + toff[i];
		const bool wrap = force_wrap || !pl.nowrap_ok || kswhost::band_clips(qlen[i], tlen[i], w);
#define RUN(C) do { if (wrap) run_one<C, true>(pl.P, qlen[i], qq, tlen[i], tt, o, c, cigar_cap); \
                    else run_one<C, false>(pl.P, qlen[i], qq, tlen[i], tt, o, c, cigar_cap); } while (0)
		switch (cpl) {
		case 2: RUN(2); break;
		case 4: RUN(4); break;
		case 8: RUN(8); break;
		case 16: RUN(16); break;
		default: return -3;
		}
	}
	return 0;
}
