// text_core.cuh -- the input and output ends of a read on the device: the original alignment out of the FASTQ comment
// (parse_ori_mapping_rst, RRH:392-429) and the SAM record of the main output (single_end_handler::output_BAM, RR:479-536,
// after htslib's sam_parse1 -> sam_format1 round trip), as host/device functions over the block's text.  One thread per read
// formats its record twice through a sink: once counting (record lengths -> prefix sums = the place of every record in the
// block's output, in input order), once writing.  The `-p` records (output_ori_bam) are rare and stay with the host.
#pragma once
#include <stdint.h>

#include "pair_core.cuh"

namespace pansvr {

struct DevRec { uint32_t name_off, comment_off, seq_off, qual_off; uint32_t name_l, comment_l, seq_l, qual_l; };   // one FASTQ record, offsets into the block text
struct StrTable { const char *pool; const uint32_t *off; uint32_t n; };   // string i = pool[off[i] .. off[i+1])
struct TextTables { StrTable target_names, sv_print, sv_id; int not_ori; };

// ---- original alignment (comment fields 1-5 and 10), numbers as atoi reads them
SEED_HD void dev_parse_ori(const char *c, uint32_t n, uint32_t seq_l, int match, DevOri &o)
{
	uint32_t tok[10], tok_l[10];
	int nt = 0;
	uint32_t i = 0;
	while (nt < 10 && i < n) {                                     // strtok_r(.., "_"): skip separators, cut at the next one
		while (i < n && c[i] == '_') ++i;
		if (i >= n) break;
		uint32_t j = i;
		while (j < n && c[j] != '_') ++j;
		tok[nt] = i; tok_l[nt] = j - i; ++nt;
		i = j + 1;
	}
	int v[5];
	for (int k = 0; k < 5; ++k) {
		long x = 0;
		if (k < nt) {
			uint32_t p = tok[k];
			const uint32_t e = p + tok_l[k];
			while (p < e && (c[p] == ' ' || c[p] == '\t')) ++p;
			bool neg = false;
			if (p < e && (c[p] == '-' || c[p] == '+')) { neg = c[p] == '-'; ++p; }
			while (p < e && c[p] >= '0' && c[p] <= '9') x = x * 10 + (c[p++] - '0');
			if (neg) x = -x;
		}
		v[k] = (int)x;
	}
	o.chr = (uint32_t)v[0]; o.ref_bg = (uint32_t)v[1]; o.read_bg = (uint32_t)v[2]; o.align_score = (uint32_t)v[3]; o.mapq = (uint8_t)v[4];
	o.direction = (nt > 9 && tok_l[9] > 0 && c[tok[9]] == 'F') ? (uint8_t)PR_FORWARD : (uint8_t)PR_REVERSE;
	o.unmapped = ((nt > 9 && tok_l[9] > 1 && c[tok[9] + 1] == 'Y') || o.chr > 24) ? 1 : 0;      // RR:413
	if (o.ref_bg >= 0x7fffffffu) o.ref_bg = 1;
	o.skip = (!o.unmapped && o.align_score == (uint32_t)((int)seq_l * match)) ? 1 : 0;      // RR:414
}

SEED_HD char dev_rev_char(char c)                                  // getReverseChar, clib/bam_file.c:320-328
{
	switch (c) { case 'A': case 'a': return 'T'; case 'C': case 'c': return 'G'; case 'G': case 'g': return 'C'; case 'T': case 't': return 'A'; default: return 'N'; }
}
SEED_HD char dev_nt16_norm(char c)                                 // seq_nt16_str[seq_nt16_table[c]]: upper case, IUPAC kept, anything else 'N'
{
	switch (c) {
	case '=': return '=';
	case 'A': case 'a': case '0': return 'A'; case 'C': case 'c': case '1': return 'C'; case 'G': case 'g': case '2': return 'G'; case 'T': case 't': case '3': return 'T';
	case 'M': case 'm': return 'M'; case 'R': case 'r': return 'R'; case 'S': case 's': return 'S'; case 'V': case 'v': return 'V';
	case 'W': case 'w': return 'W'; case 'Y': case 'y': return 'Y'; case 'H': case 'h': return 'H'; case 'K': case 'k': return 'K';
	case 'D': case 'd': return 'D'; case 'B': case 'b': return 'B'; case 'N': case 'n': return 'N';
	default: return 'N';
	}
}

// The long fields of a record (name, bases, qualities, comment) are copies with a per-byte rule:
//   COPY_PLAIN   out[i] = src[i]
//   COPY_SEQ     out[i] = nt16(src[i])                               SEQ after htslib's 4-bit round trip
//   COPY_SEQ_RC  out[i] = nt16(complement(src[len-1-i]))             getReverseStr_char, clib/bam_file.c:330-340
//   COPY_QUAL_R  out[i] = src[len-1-i], except that an even length keeps its middle pair in place: getReverseStr_qual swaps up to
//                and including i = len/2, which swaps that pair twice (clib/bam_file.c:342-360)
enum { COPY_PLAIN = 0, COPY_SEQ = 1, COPY_SEQ_RC = 2, COPY_QUAL_R = 3 };
SEED_HD char copy_byte(uint32_t mode, const char *src, uint32_t i, uint32_t len)
{
	switch (mode) {
	case COPY_SEQ: return dev_nt16_norm(src[i]);
	case COPY_SEQ_RC: return dev_nt16_norm(dev_rev_char(src[len - 1 - i]));
	case COPY_QUAL_R: return (!(len & 1) && (i == (len >> 1) - 1 || i == (len >> 1))) ? src[i] : src[len - 1 - i];
	default: return src[i];
	}
}
struct CopyJob { const char *src; char *dst; uint32_t len, mode; };

// ---- sinks: counting, writing byte by byte, and writing with the long copies set aside as jobs (the CUDA text kernel runs
// the jobs of a warp's 32 records with all lanes on one job at a time, so that its loads and stores are contiguous)
struct CountSink {
	uint32_t n;
	SEED_HD void put(char) { ++n; }
	SEED_HD void copy(const char *, uint32_t l) { n += l; }
	SEED_HD void big(const char *, uint32_t l, uint32_t) { n += l; }
	SEED_HD void patch(uint32_t) {}
	SEED_HD char *here() { return nullptr; }
};
struct WriteSink {
	char *p; uint32_t n;
	SEED_HD void put(char c) { p[n++] = c; }
	SEED_HD void copy(const char *s, uint32_t l) { for (uint32_t i = 0; i < l; ++i) p[n + i] = s[i]; n += l; }
	SEED_HD void big(const char *s, uint32_t l, uint32_t mode) { for (uint32_t i = 0; i < l; ++i) p[n + i] = copy_byte(mode, s, i, l); n += l; }
	SEED_HD void patch(uint32_t at) { p[at] = ','; }
	SEED_HD char *here() { return p + n; }
};
struct JobSink {
	char *p; uint32_t n; CopyJob job[4]; int nj; uint32_t patches[10]; int np;
	SEED_HD void put(char c) { p[n++] = c; }
	SEED_HD void copy(const char *s, uint32_t l) { for (uint32_t i = 0; i < l; ++i) p[n + i] = s[i]; n += l; }
	SEED_HD void big(const char *s, uint32_t l, uint32_t mode)
	{
		if (nj < 4) { CopyJob j; j.src = s; j.dst = p + n; j.len = l; j.mode = mode; job[nj++] = j; }
		else for (uint32_t i = 0; i < l; ++i) p[n + i] = copy_byte(mode, s, i, l);
		n += l;
	}
	SEED_HD void patch(uint32_t at) { if (np < 10) patches[np++] = at; }       // applied after the jobs have run
	SEED_HD char *here() { return p + n; }
};
template <class S> SEED_HD void put_int(S &s, long v)
{
	char b[24]; int n = 24;
	unsigned long u = v < 0 ? 0ul - (unsigned long)v : (unsigned long)v;
	do { b[--n] = (char)('0' + u % 10); u /= 10; } while (u);
	if (v < 0) b[--n] = '-';
	for (int i = n; i < 24; ++i) s.put(b[i]);
}
template <class S> SEED_HD void put_str(S &s, const StrTable &t, uint32_t i) { if (i < t.n) s.copy(t.pool + t.off[i], t.off[i + 1] - t.off[i]); else s.put('*'); }
template <class S> SEED_HD void put_lit(S &s, const char *z) { while (*z) s.put(*z++); }

// The SAM record of read `m` (0 / 1) of a pair, or nothing.  Returns 1 if a record htslib would reject was left out (bad CIGAR).
template <class S>
SEED_HD int dev_sam_record(S &s, const char *text, const DevRec &rec, const DevOri &ori, const DevFinal &f, const DevPairFinal &pf, int m,
                           const DevCand *cands, const DevCigar *cigs, const TextTables &T)
{
	if (!pf.valid || !pf.gain || !(f.flags & FIN_PRIMARY)) return 0;
	if (f.p_chr == PR_U32MAX) return 0;
	const bool is_ori = (f.flags & FIN_P_ORI) != 0;
	if (T.not_ori && is_ori) return 0;
	if (!is_ori && !(f.flags & FIN_P_CIGAR_OK)) return 1;
	const bool fwd = (f.flags & FIN_P_FWD) != 0, has_mate = (f.flags & FIN_HAS_MATE) != 0;
	const uint8_t flag = (uint8_t)((m == 0 ? 0x40 : 0) + (fwd ? 0 : 0x10) + (has_mate ? 0 : 0x08));
	s.big(text + rec.name_off, rec.name_l, COPY_PLAIN); s.put('\t'); put_int(s, flag); s.put('\t');
	put_str(s, T.target_names, f.p_chr); s.put('\t'); put_int(s, (int)f.p_ref_bg); s.put('\t'); put_int(s, (long)(f.p_mapq & 0xff)); s.put('\t');
	if (is_ori) {                                                  // the original alignment as a candidate: [S] M (RRH:421-424)
		if (ori.read_bg > 0) { put_int(s, (int16_t)(uint16_t)(int)ori.read_bg); s.put('S'); }
		put_int(s, (int16_t)(uint16_t)((int)rec.seq_l - (int)ori.read_bg)); s.put('M');
	} else {
		const DevCand &cd = cands[f.p_cand];
		if (cd.n_cig == 0) s.put('*');
		for (uint32_t k = 0; k < cd.n_cig; ++k) { const DevCigar &c = cigs[cd.cig_off + k]; put_int(s, c.size); s.put("MIDNSHP=XB"[c.type]); }
	}
	s.put('\t');
	const int isize = fwd ? pf.cur_isize : -pf.cur_isize;
	if (has_mate) {
		if (f.mate_chr == f.p_chr) s.put('='); else put_str(s, T.target_names, f.mate_chr);
		s.put('\t'); put_int(s, (int)f.mate_ref_bg); s.put('\t'); put_int(s, isize); s.put('\t');
	} else put_lit(s, "*\t0\t0\t");
	// SEQ and QUAL: reverse-complemented / reversed for a reverse-strand record (getReverseStr_char, getReverseStr_qual with its
	// double swap of the middle pair of an even length, clib/bam_file.c:330-360), SEQ through htslib's 4-bit round trip
	{
		const char *sq = text + rec.seq_off, *ql = text + rec.qual_off;
		const uint32_t L = rec.seq_l, QL = rec.qual_l, RL = rec.seq_l;   // read_l = seq.l
		if (fwd || QL == RL) {
			s.big(sq, L, fwd ? COPY_SEQ : COPY_SEQ_RC); s.put('\t'); s.big(ql, QL, fwd ? COPY_PLAIN : COPY_QUAL_R);
		} else {                                                   // (qualities of another length than the bases: the reference's in-place loops, literally)
			char *out = s.here();
			if (out) {
				for (uint32_t i = 0; i < L; ++i) out[i] = sq[i];
				out[L] = '\t';
				for (uint32_t i = 0; i < QL; ++i) out[L + 1 + i] = ql[i];
				const uint32_t half = RL >> 1;
				for (uint32_t i = 0; i < half; ++i) { const char t = out[i]; out[i] = dev_rev_char(out[RL - 1 - i]); out[RL - 1 - i] = dev_rev_char(t); }
				if (RL & 1) out[half] = dev_rev_char(out[half]);
				char *q = out + L + 1;
				for (uint32_t i = 0; i < half + 1; ++i) { const uint32_t ri = RL - 1 - i; const char t = q[i]; q[i] = q[ri]; q[ri] = t; }
				for (uint32_t i = 0; i < L; ++i) out[i] = dev_nt16_norm(out[i]);
			}
			s.n += L + 1 + QL;
		}
	}
	s.put('\t');
	put_lit(s, "AS:i:"); put_int(s, (int)f.p_align);
	put_lit(s, "\tOS:i:"); put_int(s, (int)ori.align_score);
	put_lit(s, "\tOA:Z:"); put_int(s, (int)ori.chr); s.put(','); put_int(s, (int)ori.ref_bg); s.put(','); put_int(s, (int)ori.read_bg); s.put(',');
	put_int(s, (long)ori.mapq); s.put(','); s.put(ori.unmapped ? 'U' : 'M'); s.put(';');
	if (!is_ori) { put_lit(s, "\tCS:i:"); put_int(s, (int)f.p_chain); }
	if (f.p_sv >= 0) { put_lit(s, "\tSV:Z:"); put_str(s, T.sv_print, (uint32_t)f.p_sv); }
	if (f.p_mate_sv >= 0) { put_lit(s, "\tMV:Z:"); put_str(s, T.sv_print, (uint32_t)f.p_mate_sv); }
	if (f.flags & FIN_SECONDARY) {
		put_lit(s, "\tXA:Z:"); put_int(s, (int)f.s_chr); s.put(','); put_int(s, (int)f.s_ref_bg); s.put(','); put_int(s, (int)f.s_read_bg); s.put(',');
		put_int(s, (int)f.s_align); s.put(','); s.put((f.flags & FIN_S_FWD) ? 'F' : 'R'); s.put(',');
		if (f.s_sv >= 0) put_str(s, T.sv_id, (uint32_t)f.s_sv); else s.put('*');
		s.put(';');
	}
	// RC:Z: the comment as parse_ori_mapping_rst left it: the separator after each of its first ten fields became ','
	put_lit(s, "\tRC:Z:");
	{
		const char *c = text + rec.comment_off;
		const uint32_t n = rec.comment_l, at = s.n;
		s.big(c, n, COPY_PLAIN);
		int nt = 0;
		uint32_t i = 0;
		while (nt < 10 && i < n) {
			while (i < n && c[i] == '_') ++i;
			if (i >= n) break;
			uint32_t j = i;
			while (j < n && c[j] != '_') ++j;
			++nt;
			if (j < n && j + 1 < n) s.patch(at + j);
			i = j + 1;
		}
	}
	s.put('\n');
	return 0;
}

} // namespace pansvr
