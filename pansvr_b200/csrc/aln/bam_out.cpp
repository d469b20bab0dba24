// bam_out.cpp -- see bam_out.hpp.
#include "bam_out.hpp"

#include <stdlib.h>
#include <string.h>
#include <zlib.h>
#include <algorithm>
#include <atomic>

namespace pansvr {

namespace {

enum { BGZF_BLOCK_SIZE = 0xff00, BGZF_MAX_BLOCK_SIZE = 0x10000, BLOCK_HEADER_LENGTH = 18, BLOCK_FOOTER_LENGTH = 8 };

struct Nt16 {                                      // seq_nt16_table, hts.c
	uint8_t t[256];
	Nt16()
	{
		memset(t, 15, sizeof t);
		const char *code = "=ACMGRSVTWYHKDBN";
		for (int i = 0; i < 16; ++i) { t[(unsigned char)code[i]] = (uint8_t)i; if (code[i] >= 'A') t[(unsigned char)(code[i] + 32)] = (uint8_t)i; }
		t['0'] = 1; t['1'] = 2; t['2'] = 4; t['3'] = 8;
	}
};
const Nt16 g_nt16;

int reg2bin(int64_t beg, int64_t end)              // hts_reg2bin(beg, end, 14, 5), hts.h
{
	int l, s = 14, t = ((1 << 15) - 1) / 7;
	for (--end, l = 5; l > 0; --l, s += 3, t -= 1 << ((l << 1) + l))
		if (beg >> s == end >> s) return t + (int)(beg >> s);
	return 0;
}

void put32(std::vector<uint8_t> &o, uint32_t v) { for (int i = 0; i < 4; ++i) o.push_back((uint8_t)(v >> (8 * i))); }
void put16(std::vector<uint8_t> &o, uint32_t v) { o.push_back((uint8_t)v); o.push_back((uint8_t)(v >> 8)); }
void set32(uint8_t *p, uint32_t v) { for (int i = 0; i < 4; ++i) p[i] = (uint8_t)(v >> (8 * i)); }

int cigar_op(char c)
{
	const char *ops = "MIDNSHP=XB";
	const char *p = c ? strchr(ops, c) : nullptr;
	return p ? (int)(p - ops) : -1;
}

// one tab-separated token of [p, e): returns its end (the tab or e)
const char *token_end(const char *p, const char *e) { const char *t = (const char*)memchr(p, '\t', (size_t)(e - p)); return t ? t : e; }

} // namespace

void BamHeaderInfo::parse(const std::string &header_text)
{
	text = header_text;
	names.clear(); lengths.clear(); name2id.clear();
	size_t p = 0;
	while (p < text.size()) {
		size_t e = text.find('\n', p);
		if (e == std::string::npos) e = text.size();
		if (e - p >= 3 && text.compare(p, 3, "@SQ") == 0) {
			std::string sn; uint32_t ln = 0;
			size_t q = p + 3;
			while (q < e) {                                    // tab-separated TAG:value fields
				if (text[q] != '\t') { ++q; continue; }
				const size_t fb = q + 1;
				size_t fe = text.find('\t', fb);
				if (fe == std::string::npos || fe > e) fe = e;
				if (fe - fb >= 3 && text.compare(fb, 3, "SN:") == 0) sn = text.substr(fb + 3, fe - fb - 3);
				else if (fe - fb >= 3 && text.compare(fb, 3, "LN:") == 0) ln = (uint32_t)strtoul(text.c_str() + fb + 3, nullptr, 10);
				q = fe;
			}
			if (!sn.empty()) {
				name2id[sn] = (int)names.size();                // bam_name2id: the later of two equal names wins
				names.push_back(sn); lengths.push_back(ln);
			}
		}
		p = e + 1;
	}
}

bool bam_encode_record(const char *line, size_t len, const BamHeaderInfo &h, std::vector<uint8_t> &out, std::string &err)
{
	const char *p = line, *e = line + len;
	auto fail = [&](const char *m) { err = std::string("SAM record not convertible: ") + m; return false; };
	auto name_id = [&](const char *b, const char *t) -> int {
		auto it = h.name2id.find(std::string(b, t));
		return it == h.name2id.end() ? -1 : it->second;
	};
	auto num = [&](int base, long &v) -> bool {                // strtol on a tab-terminated field
		char buf[32];
		const char *t = token_end(p, e);
		if (t == e) return false;                              // every numeric column is followed by a tab
		const size_t n = std::min<size_t>(sizeof buf - 1, (size_t)(t - p));
		memcpy(buf, p, n); buf[n] = 0;
		char *end;
		v = strtol(buf, &end, base);
		if ((size_t)(end - buf) != n) return false;            // the reference requires '\t' right after the number
		p = t + 1;
		return true;
	};
	const size_t at = out.size();
	out.resize(at + 36);                                       // block_size + 8 core words, filled at the end
	// qname
	const char *t = token_end(p, e);
	if (t == e) return fail("truncated");
	if (t - p > 10000) return fail("query name too long");
	const uint32_t l_qname = (uint32_t)(t - p) + 1;
	out.insert(out.end(), p, t); out.push_back(0);
	p = t + 1;
	long v;
	if (!num(0, v)) return fail("flag");
	uint32_t flag = (uint32_t)v & 0xffff;
	// rname
	t = token_end(p, e);
	if (t == e) return fail("truncated");
	int tid = (t - p == 1 && *p == '*') ? -1 : name_id(p, t);
	p = t + 1;
	if (!num(10, v)) return fail("pos");
	int32_t pos = (int32_t)v - 1;
	if (pos < 0 && tid >= 0) tid = -1;
	if (tid < 0) flag |= 4;
	if (!num(10, v)) return fail("mapq");
	const uint32_t mapq = (uint32_t)v & 0xff;
	// cigar
	uint32_t n_cigar = 0;
	int64_t rlen = 1, cig_qlen = 0;
	t = token_end(p, e);
	if (t == e) return fail("truncated");
	if (*p != '*') {
		for (const char *q = p; q < t; ++q) if (*q < '0' || *q > '9') ++n_cigar;
		if (n_cigar == 0) return fail("no CIGAR operations");
		int64_t ref = 0;
		const char *q = p;
		for (uint32_t i = 0; i < n_cigar; ++i) {
			long l = 0;
			bool neg = false;
			if (q < t && (*q == '-' || *q == '+')) { neg = *q == '-'; ++q; }
			while (q < t && *q >= '0' && *q <= '9') l = l * 10 + (*q++ - '0');
			if (neg) l = -l;
			const int op = q < t ? cigar_op(*q) : -1;
			if (op < 0) return fail("unrecognized CIGAR operator");
			++q;
			const uint32_t w = (uint32_t)l << 4 | (uint32_t)op;
			put32(out, w);
			const uint32_t ol = w >> 4;
			if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) ref += ol;       // bam_cigar2rlen
			if (op == 0 || op == 1 || op == 4 || op == 7 || op == 8) cig_qlen += ol;  // bam_cigar2qlen
		}
		rlen = (flag & 4) ? 1 : ref;
	} else flag |= 4;
	p = t + 1;
	const uint32_t bin = (uint32_t)reg2bin(pos, (int64_t)pos + rlen) & 0xffff;
	// mate
	t = token_end(p, e);
	if (t == e) return fail("truncated");
	int mtid;
	if (t - p == 1 && *p == '=') mtid = tid;
	else if (t - p == 1 && *p == '*') mtid = -1;
	else mtid = name_id(p, t);
	p = t + 1;
	if (!num(10, v)) return fail("mpos");
	const int32_t mpos = (int32_t)v - 1;
	if (mpos < 0 && mtid >= 0) mtid = -1;
	if (!num(10, v)) return fail("tlen");
	const int32_t isize = (int32_t)v;
	// seq
	t = token_end(p, e);
	if (t == e) return fail("truncated");
	uint32_t l_qseq = 0;
	if (!(t - p == 1 && *p == '*')) {
		l_qseq = (uint32_t)(t - p);
		if (n_cigar && cig_qlen != (int64_t)l_qseq) return fail("CIGAR and query sequence are of different length");
		const size_t s0 = out.size();
		out.resize(s0 + (l_qseq + 1) / 2, 0);
		for (uint32_t i = 0; i < l_qseq; ++i) out[s0 + (i >> 1)] |= (uint8_t)(g_nt16.t[(unsigned char)p[i]] << ((~i & 1) << 2));
	}
	p = t + 1;
	// qual (the last mandatory column: may end the line)
	t = token_end(p, e);
	if (!(t - p == 1 && *p == '*')) {
		if ((uint32_t)(t - p) != l_qseq) return fail("SEQ and QUAL are of different length");
		for (uint32_t i = 0; i < l_qseq; ++i) out.push_back((uint8_t)(p[i] - 33));
	} else out.insert(out.end(), l_qseq, 0xff);
	p = t < e ? t + 1 : e;
	// aux
	while (p < e) {
		t = token_end(p, e);
		if (t - p < 5) return fail("incomplete aux field");
		out.push_back((uint8_t)p[0]); out.push_back((uint8_t)p[1]);
		const char type = p[3];
		const char *q = p + 5;
		if (type != 'Z' && type != 'H' && t - q < 1) return fail("incomplete aux field");
		std::string val(q, t);                                 // NUL-terminated copy for the strto* calls
		if (type == 'A' || type == 'a' || type == 'c' || type == 'C') { out.push_back('A'); out.push_back((uint8_t)*q); }
		else if (type == 'i' || type == 'I') {
			if (*q == '-') {
				const long x = strtol(val.c_str(), nullptr, 10);
				if (x >= -128) { out.push_back('c'); out.push_back((uint8_t)x); }
				else if (x >= -32768) { out.push_back('s'); put16(out, (uint32_t)x); }
				else { out.push_back('i'); put32(out, (uint32_t)x); }
			} else {
				const unsigned long x = strtoul(val.c_str(), nullptr, 10);
				if (x <= 0xff) { out.push_back('C'); out.push_back((uint8_t)x); }
				else if (x <= 0xffff) { out.push_back('S'); put16(out, (uint32_t)x); }
				else { out.push_back('I'); put32(out, (uint32_t)x); }
			}
		} else if (type == 'f') {
			const float f = (float)strtod(val.c_str(), nullptr);
			uint32_t w; memcpy(&w, &f, 4);
			out.push_back('f'); put32(out, w);
		} else if (type == 'd') {
			const double d = strtod(val.c_str(), nullptr);
			uint64_t w; memcpy(&w, &d, 8);
			out.push_back('d'); put32(out, (uint32_t)w); put32(out, (uint32_t)(w >> 32));
		} else if (type == 'Z' || type == 'H') {
			if (type == 'H' && (val.size() & 1)) return fail("hex field does not have an even number of digits");
			out.push_back((uint8_t)type); out.insert(out.end(), val.begin(), val.end()); out.push_back(0);
		} else if (type == 'B') {
			if (val.size() < 2) return fail("incomplete B-typed aux field");
			const char sub = val[0];
			const int size = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : (sub == 'i' || sub == 'I' || sub == 'f') ? 4 : 0;
			if (!size) return fail("unrecognized type B");
			if (val.size() > 1 && val[1] != ',') return fail("B aux field type not followed by ','");
			uint32_t n = 0;
			for (size_t i = 1; i < val.size(); ++i) n += val[i] == ',';
			out.push_back('B'); out.push_back((uint8_t)sub); put32(out, n);
			const char *r = val.c_str() + 1, *re = val.c_str() + val.size();
			while (r + 1 < re) {                               // r at a ','
				char *end;
				if (sub == 'f') { const float f = strtof(r + 1, &end); uint32_t w; memcpy(&w, &f, 4); put32(out, w); }
				else if (sub == 'c' || sub == 's' || sub == 'i') { const long x = strtol(r + 1, &end, 0); if (size == 1) out.push_back((uint8_t)x); else if (size == 2) put16(out, (uint32_t)x); else put32(out, (uint32_t)x); }
				else { const unsigned long x = strtoul(r + 1, &end, 0); if (size == 1) out.push_back((uint8_t)x); else if (size == 2) put16(out, (uint32_t)x); else put32(out, (uint32_t)x); }
				r = end;
				while (r < re && *r != ',') ++r;
			}
		} else return fail("unrecognized aux type");
		p = t < e ? t + 1 : e;
	}
	const uint32_t block_len = (uint32_t)(out.size() - at - 4);
	uint8_t *c = out.data() + at;
	set32(c, block_len);
	set32(c + 4, (uint32_t)tid);
	set32(c + 8, (uint32_t)pos);
	set32(c + 12, bin << 16 | mapq << 8 | (l_qname & 0xff));
	set32(c + 16, flag << 16 | (n_cigar & 0xffff));
	set32(c + 20, l_qseq);
	set32(c + 24, (uint32_t)mtid);
	set32(c + 28, (uint32_t)mpos);
	set32(c + 32, (uint32_t)isize);
	return true;
}

// ---- BGZF
namespace {

// one BGZF block of `slen` bytes (bgzf_compress, zlib path): header, raw deflate at the default level, CRC32 + ISIZE
bool bgzf_block(const uint8_t *src, uint32_t slen, std::vector<uint8_t> &dst)
{
	static const uint8_t magic[BLOCK_HEADER_LENGTH] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0, 0};
	dst.resize(BGZF_MAX_BLOCK_SIZE);
	z_stream zs;
	memset(&zs, 0, sizeof zs);
	zs.next_in = (Bytef*)src; zs.avail_in = slen;
	zs.next_out = dst.data() + BLOCK_HEADER_LENGTH; zs.avail_out = BGZF_MAX_BLOCK_SIZE - BLOCK_HEADER_LENGTH - BLOCK_FOOTER_LENGTH;
	if (deflateInit2(&zs, Z_DEFAULT_COMPRESSION, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
	if (deflate(&zs, Z_FINISH) != Z_STREAM_END) { deflateEnd(&zs); return false; }
	const size_t dlen = zs.total_out + BLOCK_HEADER_LENGTH + BLOCK_FOOTER_LENGTH;
	deflateEnd(&zs);
	memcpy(dst.data(), magic, BLOCK_HEADER_LENGTH);
	dst[16] = (uint8_t)((dlen - 1) & 0xff); dst[17] = (uint8_t)((dlen - 1) >> 8);
	const uint32_t crc = (uint32_t)crc32(crc32(0L, nullptr, 0), src, slen);
	set32(dst.data() + dlen - 8, crc);
	set32(dst.data() + dlen - 4, slen);
	dst.resize(dlen);
	return true;
}

} // namespace

bool BamWriter::flush_blocks(const uint8_t *data, const std::vector<std::pair<size_t, uint32_t>> &blocks)
{
	const size_t nb = blocks.size();
	if (nb == 0) return true;
	std::vector<std::vector<uint8_t>> comp(nb);
	std::atomic<bool> ok(true);
	par_(nb, [&](size_t b, size_t e, int) {
		for (size_t i = b; i < e; ++i) if (!bgzf_block(data + blocks[i].first, blocks[i].second, comp[i])) ok = false;
	});
	if (!ok) return false;
	for (size_t i = 0; i < nb; ++i) if (fwrite(comp[i].data(), 1, comp[i].size(), f_) != comp[i].size()) return false;
	return true;
}

bool BamWriter::write_header(const BamHeaderInfo &h)
{
	std::vector<uint8_t> b;
	b.insert(b.end(), {'B', 'A', 'M', 1});
	put32(b, (uint32_t)h.text.size());
	b.insert(b.end(), h.text.begin(), h.text.end());
	put32(b, (uint32_t)h.names.size());
	for (size_t i = 0; i < h.names.size(); ++i) {
		put32(b, (uint32_t)h.names[i].size() + 1);
		b.insert(b.end(), h.names[i].begin(), h.names[i].end()); b.push_back(0);
		put32(b, h.lengths[i]);
	}
	// bgzf_write fills blocks of BGZF_BLOCK_SIZE; bam_hdr_write ends with bgzf_flush
	std::vector<std::pair<size_t, uint32_t>> blocks;
	for (size_t off = 0; off < b.size(); off += BGZF_BLOCK_SIZE) blocks.push_back(std::make_pair(off, (uint32_t)std::min<size_t>(BGZF_BLOCK_SIZE, b.size() - off)));
	return flush_blocks(b.data(), blocks);
}

bool BamWriter::write_records(const uint8_t *recs, const std::vector<uint32_t> &sizes)
{
	// The byte stream is pending_ followed by recs; replay bgzf's block_offset over the record sizes to find the cuts.
	std::vector<uint8_t> head;                                 // blocks that start inside pending_ are assembled here
	std::vector<std::pair<size_t, uint32_t>> blocks;           // (offset into recs, length) of blocks made of recs bytes only
	size_t block_start = 0;                                    // offset in recs where the open block's recs part starts
	size_t block_offset = pending_.size();                     // bytes in the open block
	size_t cur = 0;                                            // offset in recs
	bool head_open = !pending_.empty();                        // the open block still begins with pending_
	std::vector<std::vector<uint8_t>> head_blocks;
	auto cut = [&](size_t upto) {                              // close the open block at recs offset `upto`
		if (block_offset == 0) return;
		if (head_open) {
			std::vector<uint8_t> hb(pending_);
			hb.insert(hb.end(), recs + block_start, recs + upto);
			head_blocks.push_back(std::move(hb));
			pending_.clear(); head_open = false;
		} else blocks.push_back(std::make_pair(block_start, (uint32_t)(upto - block_start)));
		block_start = upto; block_offset = 0;
	};
	for (uint32_t sz : sizes) {
		if (block_offset + sz > BGZF_BLOCK_SIZE) cut(cur);      // bgzf_flush_try(fp, 4 + block_len)
		size_t remaining = sz;
		while (remaining > 0) {                                 // bgzf_write
			const size_t copy = std::min<size_t>(BGZF_BLOCK_SIZE - block_offset, remaining);
			block_offset += copy; cur += copy; remaining -= copy;
			if (block_offset == BGZF_BLOCK_SIZE) cut(cur);
		}
	}
	// compress: the (at most one) block that began in pending_, then the rest
	for (std::vector<uint8_t> &hb : head_blocks) {
		std::vector<std::pair<size_t, uint32_t>> one(1, std::make_pair((size_t)0, (uint32_t)hb.size()));
		if (!flush_blocks(hb.data(), one)) return false;
	}
	if (!flush_blocks(recs, blocks)) return false;
	if (head_open) pending_.insert(pending_.end(), recs + block_start, recs + cur);   // everything still fits the open block
	else pending_.assign(recs + block_start, recs + cur);
	return true;
}

bool BamWriter::close()
{
	if (!pending_.empty()) {
		std::vector<std::pair<size_t, uint32_t>> one(1, std::make_pair((size_t)0, (uint32_t)pending_.size()));
		if (!flush_blocks(pending_.data(), one)) return false;
		pending_.clear();
	}
	std::vector<uint8_t> eof;
	if (!bgzf_block((const uint8_t*)"", 0, eof)) return false;
	return fwrite(eof.data(), 1, eof.size(), f_) == eof.size() && fflush(f_) == 0;
}

} // namespace pansvr
