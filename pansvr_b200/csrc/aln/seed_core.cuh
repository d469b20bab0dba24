// seed_core.cuh -- k-mer lookup and in-unipath MEM extension of one read strand.
//
// Same results, in the same order, as the reference's seeding loop
//   single_end_handler::chainning_one_read   read_realignment.cpp:614-635
//   deBGA_INDEX::search_kmer                 deBGA_index.cpp:84-101   (+ binsearch_range, clib/binarys_qsort.c:4-80)
//   deBGA_INDEX::UNITIG_MEM_search           deBGA_index.cpp:105-146  (+ binsearch_interval_unipath64, clib/binarys_qsort.c:162)
// written as host/device functions over a view of the index arrays.  The product runs them in
// seed_kernel (seed_gpu.cu), one thread per read strand against the device-resident index; the
// CPU-only test build (tests/emul) steps the same functions on the host.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SEED_HD __host__ __device__ __forceinline__
#else
#define SEED_HD inline
#endif

namespace pansvr {

enum { LEN_KMER = 20, SEED_STEP = 5, UNI_POS_N_MAX = 32, SEED_K_OFF = 4 /* (K_T - LEN_KMER) << 1 */ };

struct IndexView {                 // device or host pointers to the index arrays
	const uint64_t *seqb;          // 2-bit unipath sequence, 32 bases per word, MSB first
	const uint64_t *seqf;          // start offset of each unipath in seqb
	const uint64_t *posp;          // CSR pointers into the reference-position list
	const uint32_t *bkt_dir;       // compacted bucket table (index.hpp): directory over the top 20 bits of the bucket number,
	const uint32_t *bkt_key;       //   non-empty bucket numbers,
	const uint64_t *bkt_start;     //   and their starts (hash_g[h]); lookups give exactly hash_g[h], hash_g[h+1]
	const uint64_t *off_g;         // offset of each indexed 22-mer in seqb
	const uint32_t *kmer_g;        // low 16 bits of each indexed 22-mer
	uint64_t n_seqf;
};

struct Mem {                       // vertex_MEM, deBGA_index.hpp:23-52
	uint64_t uid;
	uint32_t seed_id, read_pos, uni_pos_off, length, pos_n, pad_;
};

// 20-mer starting at read_off of a read packed 32 bases per word, MSB first (read_realignment.cpp:204-210)
SEED_HD uint64_t get_kmer(uint32_t read_off, const uint64_t *read_bit)
{
	const uint32_t w = read_off >> 5, k = read_off & 0x1f;
	const uint64_t full = (read_bit[w] << (k << 1)) | (k == 0 ? 0 : (read_bit[w + 1] >> ((32 - k) << 1)));
	return full >> ((32 - LEN_KMER) << 1);
}

SEED_HD uint32_t base_at(const uint64_t *seq, uint64_t i) { return (uint32_t)(seq[i >> 5] >> ((31 - (i & 0x1f)) << 1)) & 3u; }

// range of index k-mers equal to the read 20-mer: 14 bases through the (compacted) bucket table, the remaining 6 against
// kmer_g >> 4 by binary search for the first and last equal key.  false = no hit.
SEED_HD bool search_kmer(const IndexView &ix, uint64_t kmer, int64_t range[2])
{
	const uint64_t key = kmer & 0xfff, h = kmer >> 12;
	uint32_t blo = ix.bkt_dir[h >> 8];
	const uint32_t bend = ix.bkt_dir[(h >> 8) + 1];
	uint32_t bhi = bend;
	while (blo < bhi) {                                // first non-empty bucket >= h among those sharing h's top 20 bits
		const uint32_t m = (blo + bhi) >> 1;
		if (ix.bkt_key[m] < (uint32_t)h) blo = m + 1; else bhi = m;
	}
	if (blo == bend || ix.bkt_key[blo] != (uint32_t)h) return false;   // hash_g[h+1] == hash_g[h]: empty bucket
	const uint64_t base = ix.bkt_start[blo];
	const int64_t n = (int64_t)(ix.bkt_start[blo + 1] - base);
	const uint32_t *v = ix.kmer_g + base;
	int64_t l = 0, r = n - 1;
	while (l <= r) {
		const int64_t m = (l + r) / 2;
		const uint32_t tmp = v[m] >> SEED_K_OFF;
		if (tmp == key) {
			range[0] = range[1] = m;
			int64_t sl = l, sr = m - 1;
			while (sl <= sr) {                         // lowest equal key
				const int64_t sm = (sl + sr) / 2;
				const uint32_t t = v[sm] >> SEED_K_OFF;
				if (t == key) { range[0] = sm; sr = sm - 1; }
				else if (t > key) sr = sm - 1;
				else sl = sm + 1;
			}
			sl = m + 1; sr = r;
			while (sl <= sr) {                         // highest equal key
				const int64_t sm = (sl + sr) / 2;
				const uint32_t t = v[sm] >> SEED_K_OFF;
				if (t == key) { range[1] = sm; sl = sm + 1; }
				else if (t > key) sr = sm - 1;
				else sl = sm + 1;
			}
			range[0] += (int64_t)base; range[1] += (int64_t)base;
			return true;
		} else if (tmp > key) r = m - 1;
		else l = m + 1;
	}
	return false;
}

// unipath that contains offset x: last start <= x
SEED_HD int64_t unipath_of(const IndexView &ix, uint64_t x)
{
	int64_t low = 0, high = (int64_t)ix.n_seqf - 1;
	while (low <= high) {
		const int64_t mid = (low + high) >> 1;
		if (x < ix.seqf[mid]) high = mid - 1;
		else if (x > ix.seqf[mid]) low = mid + 1;
		else return mid;
	}
	return high;
}

// ---- word-wise comparison of packed sequences (32 bases per 64-bit word, MSB first).  Both arrays carry one spare word
// after their last base (index.cpp pads unipath.seqb, stage A pads every packed read), so a window may read one word past
// the last base; what it sees there is cut off by the caller's cap.
SEED_HD uint32_t clz64(uint64_t x)
{
#if defined(__CUDA_ARCH__)
	return (uint32_t)__clzll((long long)x);
#else
	return (uint32_t)__builtin_clzll(x);
#endif
}
SEED_HD uint32_t ctz64(uint64_t x)
{
#if defined(__CUDA_ARCH__)
	return (uint32_t)(__ffsll((long long)x) - 1);
#else
	return (uint32_t)__builtin_ctzll(x);
#endif
}
// the 32 bases pos .. pos+31
SEED_HD uint64_t window_at(const uint64_t *seq, uint64_t pos)
{
	const uint64_t w = pos >> 5;
	const uint32_t k = (uint32_t)(pos & 0x1f);
	return k ? (seq[w] << (k << 1)) | (seq[w + 1] >> ((32 - k) << 1)) : seq[w];
}
// the 32 bases pos-32 .. pos-1 (pos >= 1); for pos < 32 the pos existing bases, right-aligned
SEED_HD uint64_t window_before(const uint64_t *seq, uint64_t pos)
{
	return pos >= 32 ? window_at(seq, pos - 32) : seq[0] >> ((32 - (uint32_t)pos) << 1);
}
// length of the common prefix of a[ap..] and b[bp..], at most cap
SEED_HD uint32_t match_right(const uint64_t *a, uint64_t ap, const uint64_t *b, uint64_t bp, uint32_t cap)
{
	uint32_t m = 0;
	while (m < cap) {
		const uint64_t x = window_at(a, ap + m) ^ window_at(b, bp + m);
		if (x) { m += clz64(x) >> 1; break; }
		m += 32;
	}
	return m < cap ? m : cap;
}
// length of the common suffix of a[..ap) and b[..bp), at most cap (cap <= ap, cap <= bp)
SEED_HD uint32_t match_left(const uint64_t *a, uint64_t ap, const uint64_t *b, uint64_t bp, uint32_t cap)
{
	uint32_t m = 0;
	while (m < cap) {
		const uint64_t x = window_before(a, ap - m) ^ window_before(b, bp - m);
		if (x) { m += ctz64(x) >> 1; break; }
		m += 32;
	}
	return m < cap ? m : cap;
}

// extend one k-mer hit to a maximal exact match inside its unipath.  The reference walks base by base
// (deBGA_index.cpp:118-131: left_i / right_i stop at the first mismatch, at the unipath end or at the read end);
// the same counts come out of 32-base XOR windows.
SEED_HD void unitig_mem(const IndexView &ix, uint64_t kmer_index, const uint64_t *read_bit, uint32_t read_off, uint32_t read_length,
                        uint32_t seed_id, Mem &out, uint32_t &max_right_i)
{
	const uint64_t kpos = ix.off_g[kmer_index];
	const int64_t uid = unipath_of(ix, kpos);
	const uint32_t off_l = (uint32_t)(kpos - ix.seqf[uid]);
	const uint32_t off_r = (uint32_t)(ix.seqf[uid + 1] - (kpos + LEN_KMER));
	const uint32_t cap_l = off_l < read_off ? off_l : read_off;
	const uint32_t room_r = read_length - read_off - LEN_KMER;
	const uint32_t cap_r = off_r < room_r ? off_r : room_r;
	const uint32_t left_i = 1 + match_left(ix.seqb, kpos, read_bit, read_off, cap_l);
	const uint32_t right_i = 1 + match_right(ix.seqb, kpos + LEN_KMER, read_bit, read_off + LEN_KMER, cap_r);
	out.uid = (uint64_t)uid;
	out.seed_id = seed_id;
	out.read_pos = read_off + 1 - left_i;
	out.uni_pos_off = off_l + 1 - left_i;
	out.length = LEN_KMER + left_i + right_i - 2;
	out.pos_n = (uint32_t)(ix.posp[uid + 1] - ix.posp[uid]);
	out.pad_ = 0;
	if (right_i > max_right_i) max_right_i = right_i;
}

// All MEMs of one read strand, in the reference's order.  seed_list (only read when is_str) marks the seed
// positions an STR read may use.  Returns the number of MEMs found; only the first `cap` are stored.  *n_probes (optional)
// receives the number of k-mer lookups made (the unit of the seeding roofline).
SEED_HD int seed_read_strand(const IndexView &ix, const uint64_t *read_bit, uint32_t read_l, bool is_str, const uint8_t *seed_list,
                             Mem *out, int cap, uint32_t *n_probes = nullptr)
{
	int n = 0;
	uint32_t probes = 0;
	const uint32_t kmer_number = read_l - LEN_KMER + 1;
	uint32_t max_search_right = 0;
	for (uint32_t read_off = 0; read_off < kmer_number; read_off += SEED_STEP) {
		if (read_off + LEN_KMER - 1 <= max_search_right) continue;          // still inside the last MEM
		if (is_str && seed_list[read_off] == 0) continue;
		int64_t range[2];
		++probes;
		if (!search_kmer(ix, get_kmer(read_off, read_bit), range)) continue;
		if ((uint64_t)(range[1] - range[0] + 1) > UNI_POS_N_MAX) continue;
		uint32_t max_right_i = 1;
		for (int64_t hit = range[0]; hit <= range[1]; ++hit) {
			Mem m;
			unitig_mem(ix, (uint64_t)hit, read_bit, read_off, read_l, (uint32_t)n, m, max_right_i);
			if (n < cap) out[n] = m;
			++n;
		}
		max_search_right = read_off + LEN_KMER + max_right_i - 1;
	}
	if (n_probes) *n_probes = probes;
	return n;
}

} // namespace pansvr
