// emul_ksw.cpp -- TEST INFRASTRUCTURE ONLY.
// Compiles the warp program of pansvr_b200/csrc/ksw_team.cuh for the host (PANSVR_HOST_EMUL) and
// steps it on the 32-fibre lock-step warp of warp_emul.hpp, so the kernel's logic can be checked
// against the oracle on a machine without a GPU.  Never linked into the product library.
#define PANSVR_HOST_EMUL 1
#include <vector>
#include <string.h>
#include "../../pansvr_b200/csrc/ksw_host.hpp"
#include "../../pansvr_b200/csrc/ksw_team.cuh"

// ---- team kernel: 32/TEAM alignments per emulated warp
template <int TEAM, bool WRAP, bool WC>
static void run_team_wc(const kswfast::Params &P, int nt, const int *ids, const uint8_t *qseq, const int64_t *qoff, const int32_t *qlen,
                     const uint8_t *tseq, const int64_t *toff, const int32_t *tlen, int32_t *res, uint32_t *cigar, int cigar_cap)
{
	constexpr int NT = 32 / TEAM, W = 16 * TEAM;
	int maxq = 0, maxrows = 0;
	for (int k = 0; k < nt; ++k) {
		maxq = std::max(maxq, qlen[ids[k]]);
		maxrows = std::max(maxrows, kswhost::n_diagonals(qlen[ids[k]], tlen[ids[k]], P.w));
	}
	const int per_team = kswteam::team_smem_bytes(TEAM, maxq);
	std::vector<uint8_t> smem((size_t)per_team * NT + 64, 0xCD);
	std::vector<uint32_t> scr(32 * 8, 0xDEADBEEF), mtab(128);
	kswteam::fill_mask_table(mtab.data(), 0, 1);
	const size_t tb_per_team = ((size_t)maxrows + 1) * W + 64;
	std::vector<uint8_t> tb(tb_per_team * NT, 0xEE);
	uint8_t *sm = (uint8_t*)(((uintptr_t)smem.data() + 15) & ~(uintptr_t)15);
	WarpEmul::run([&]() {
		const int lane = WarpEmul::lane(), team = lane / TEAM;
		const bool have = team < nt;
		const int id = have ? ids[team] : 0;
		uint8_t *base = sm + (size_t)team * per_team;
		int32_t *Hs = (int32_t*)base, *Hsnap = Hs + W;
		uint8_t *QS = (uint8_t*)(Hsnap + W);
		uint8_t *Ssp = base + per_team - 32;
		kswteam::align_team<TEAM, WRAP, WC>(P, have, have ? qlen[id] : 0, qseq + (have ? qoff[id] : 0), have ? tlen[id] : 0,
		                                tseq + (have ? toff[id] : 0), res + (size_t)id * kswfast::RES_WORDS,
		                                cigar + (size_t)id * cigar_cap, cigar_cap, tb.data() + (size_t)team * tb_per_team, Hs, Hsnap,
		                                QS, Ssp, scr.data() + lane * 8, mtab.data());
	});
}

template <int TEAM, bool WRAP>
static void run_team(const kswfast::Params &P, int nt, const int *ids, const uint8_t *qseq, const int64_t *qoff, const int32_t *qlen,
                     const uint8_t *tseq, const int64_t *toff, const int32_t *tlen, int32_t *res, uint32_t *cigar, int cigar_cap)
{
	if (P.flag & kswfast::F_SCORE_ONLY) run_team_wc<TEAM, WRAP, false>(P, nt, ids, qseq, qoff, qlen, tseq, toff, tlen, res, cigar, cigar_cap);
	else run_team_wc<TEAM, WRAP, true>(P, nt, ids, qseq, qoff, qlen, tseq, toff, tlen, res, cigar, cigar_cap);
}

extern "C" int emul_ksw_team_batch(int n, const uint8_t *qseq, const int64_t *qoff, const int32_t *qlen, const uint8_t *tseq,
                                   const int64_t *toff, const int32_t *tlen, int m, const int8_t *mat, int q, int e, int q2,
                                   int e2, int w, int zdrop, int end_bonus, int flag, int32_t *res, uint32_t *cigar,
                                   int cigar_cap, int force_team, int force_wrap)
{
	kswhost::Plan pl = kswhost::make_plan(m, mat, q, e, q2, e2, w, zdrop, end_bonus, flag);
	if (!pl.trivial && !pl.fast_params) return -1;
	// group tasks by (team size, wrap) in input order, run them 32/TEAM at a time
	std::vector<int> bucket[6][2];
	for (int i = 0; i < n; ++i) {
		int32_t *o = res + (size_t)i * kswfast::RES_WORDS;
		memset(cigar + (size_t)i * cigar_cap, 0, sizeof(uint32_t) * cigar_cap);
		if (pl.trivial || qlen[i] <= 0 || tlen[i] <= 0) {
			const int32_t z[12] = {0, 0, -1, -1, kswfast::NEG_INF, -1, kswfast::NEG_INF, -1, kswfast::NEG_INF, 0, 0, 0};
			memcpy(o, z, sizeof(z));
			continue;
		}
		int team = kswhost::pick_team(qlen[i], tlen[i], w);
		if (team == 0) return -2;
		if (force_team > team) team = force_team;
		const bool wrap = force_wrap || !pl.nowrap_ok || kswhost::band_clips(qlen[i], tlen[i], w);
		int lg = 0; while ((2 << lg) < team) ++lg;
		bucket[lg][wrap].push_back(i);
	}
	for (int lg = 0; lg < 5; ++lg)
		for (int wr = 0; wr < 2; ++wr) {
			const std::vector<int> &b = bucket[lg][wr];
			const int team = 2 << lg, nt = 32 / team;
			for (size_t s0 = 0; s0 < b.size(); s0 += nt) {
				const int cnt = (int)std::min<size_t>(nt, b.size() - s0);
#define RUNT(T) do { if (wr) run_team<T, true>(pl.P, cnt, &b[s0], qseq, qoff, qlen, tseq, toff, tlen, res, cigar, cigar_cap); \
                     else run_team<T, false>(pl.P, cnt, &b[s0], qseq, qoff, qlen, tseq, toff, tlen, res, cigar, cigar_cap); } while (0)
				switch (team) { case 2: RUNT(2); break; case 4: RUNT(4); break; case 8: RUNT(8); break; case 16: RUNT(16); break; default: RUNT(32); }
			}
		}
	return 0;
}
