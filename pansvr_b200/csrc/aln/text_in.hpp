// text_in.hpp -- the FASTQ input of the command line as a byte stream: plain text, gzip, or BGZF (what htslib's bgzip writes).
//
// The reference reads its input through zlib's gzread (kseq.h over gzFile, clib/utils.c:953-990): one inflate stream on one
// thread.  A gzip member has no index, so a plain .gz stays that way here too (zlib's gzread, as in the reference).  A BGZF
// file is a series of independent gzip members of at most 64 KiB, each carrying its compressed size in an extra field
// ('B','C'; SAM specification 4.1): those are read a batch at a time and inflated on all helper threads, so that
// decompression keeps up with the device path.  The bytes delivered are the same either way.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <zlib.h>
#include <string>
#include <thread>
#include <vector>

namespace pansvr {

class TextInput {
public:
	// path "-" = standard input.  threads: how many threads inflate BGZF blocks.
	bool open(const char *path, int threads, std::string &err)
	{
		threads_ = threads < 1 ? 1 : threads;
		const bool is_stdin = strcmp(path, "-") == 0;
		f_ = is_stdin ? stdin : fopen(path, "rb");
		if (!f_) { err = std::string("cannot open ") + path; return false; }
		own_ = !is_stdin;
		// look at the first member's header; what was read is kept and handed to whichever reader takes over (a pipe cannot rewind)
		uint8_t h[18];
		const size_t got = fread(h, 1, sizeof h, f_);
		bgzf_ = got == sizeof h && is_bgzf_header(h);
		if (bgzf_) { pend_.assign(h, h + got); return true; }
		if (!is_stdin) {                                          // plain text or gzip in a file: zlib's own reader, like the reference
			fclose(f_); f_ = nullptr;
			gz_ = gzopen(path, "r");
		} else {
			// a pipe that is not BGZF: zlib must see the bytes already taken.  gzdopen cannot be given a prefix, so the stream is
			// inflated here with inflate() for gzip, or passed through for plain text.
			pend_.assign(h, h + got);
			raw_ = true;
			raw_gzip_ = got >= 2 && h[0] == 0x1f && h[1] == 0x8b;
			if (raw_gzip_) {
				memset(&zs_, 0, sizeof zs_);
				if (inflateInit2(&zs_, 15 + 16) != Z_OK) { err = "zlib initialisation failed"; return false; }
				zs_open_ = true;
				cbuf_.resize(1 << 20);
			}
			return true;
		}
		if (!gz_) { err = std::string("cannot open ") + path; return false; }
		gzbuffer(gz_, 1 << 20);
		return true;
	}
	bool is_bgzf() const { return bgzf_; }
	// up to n bytes of text; 0 = end of the input, -1 = error (why() says what)
	long read(char *dst, size_t n)
	{
		if (gz_) { const int got = gzread(gz_, dst, (unsigned)std::min<size_t>(n, 1u << 30)); if (got < 0) why_ = "gzread failed"; return got; }
		if (raw_) return raw_read(dst, n);
		size_t out = 0;
		while (out < n) {
			if (text_at_ == text_.size()) { if (!next_batch()) { if (!why_.empty()) return -1; break; } continue; }
			const size_t take = std::min(n - out, text_.size() - text_at_);
			memcpy(dst + out, text_.data() + text_at_, take);
			text_at_ += take; out += take;
		}
		return (long)out;
	}
	const std::string &why() const { return why_; }
	void close()
	{
		if (gz_) { gzclose(gz_); gz_ = nullptr; }
		if (zs_open_) { inflateEnd(&zs_); zs_open_ = false; }
		if (f_ && own_) fclose(f_);
		f_ = nullptr;
	}
	~TextInput() { close(); }

private:
	static bool is_bgzf_header(const uint8_t *h)
	{
		return h[0] == 0x1f && h[1] == 0x8b && h[2] == 8 && (h[3] & 4) && h[10] == 6 && h[11] == 0 && h[12] == 'B' && h[13] == 'C' && h[14] == 2 && h[15] == 0;
	}
	size_t fill(uint8_t *dst, size_t n)                           // from the bytes already taken, then the file
	{
		size_t got = 0;
		if (pend_at_ < pend_.size()) { got = std::min(n, pend_.size() - pend_at_); memcpy(dst, pend_.data() + pend_at_, got); pend_at_ += got; }
		if (got < n) got += fread(dst + got, 1, n - got, f_);
		return got;
	}
	long raw_read(char *dst, size_t n)
	{
		if (!raw_gzip_) return (long)fill((uint8_t*)dst, n);
		zs_.next_out = (Bytef*)dst; zs_.avail_out = (uInt)std::min<size_t>(n, 1u << 30);
		const uInt want = zs_.avail_out;
		while (zs_.avail_out == want) {
			if (zs_.avail_in == 0) {
				const size_t got = fill(cbuf_.data(), cbuf_.size());
				if (got == 0) break;
				zs_.next_in = cbuf_.data(); zs_.avail_in = (uInt)got;
			}
			const int rc = inflate(&zs_, Z_NO_FLUSH);
			if (rc == Z_STREAM_END) { if (inflateReset(&zs_) != Z_OK) { why_ = "zlib reset failed"; return -1; } continue; }   // next member, if any
			if (rc != Z_OK && rc != Z_BUF_ERROR) { why_ = "the gzip stream is damaged"; return -1; }
			if (rc == Z_BUF_ERROR && zs_.avail_in == 0) continue;
		}
		return (long)(want - zs_.avail_out);
	}
	// reads up to BATCH blocks and inflates them on all threads; text_ = their text in order.  false = nothing more (or an error)
	bool next_batch()
	{
		enum { BATCH = 512 };
		text_.clear(); text_at_ = 0;
		if (eof_) return false;
		struct Block { size_t c_off, c_len, t_off, t_len; uint32_t crc; };
		std::vector<Block> blocks;
		comp_.clear();
		size_t t_total = 0;
		while (blocks.size() < BATCH) {
			uint8_t h[18];
			const size_t got = fill(h, sizeof h);
			if (got == 0) { eof_ = true; break; }
			if (got != sizeof h || !is_bgzf_header(h)) { why_ = "the BGZF input is damaged (a block header was expected)"; return false; }
			const size_t bsize = (size_t)(h[16] | h[17] << 8) + 1;
			if (bsize < 18 + 8) { why_ = "the BGZF input is damaged (block size)"; return false; }
			const size_t rest = bsize - 18, at = comp_.size();
			comp_.resize(at + rest);
			if (fill(comp_.data() + at, rest) != rest) { why_ = "the BGZF input ends inside a block"; return false; }
			const uint8_t *tail = comp_.data() + at + rest - 8;
			Block b;
			b.c_off = at; b.c_len = rest - 8;
			b.crc = (uint32_t)tail[0] | (uint32_t)tail[1] << 8 | (uint32_t)tail[2] << 16 | (uint32_t)tail[3] << 24;
			b.t_len = (size_t)tail[4] | (size_t)tail[5] << 8 | (size_t)tail[6] << 16 | (size_t)tail[7] << 24;
			if (b.t_len > 65536) { why_ = "the BGZF input is damaged (block text size)"; return false; }
			b.t_off = t_total; t_total += b.t_len;
			blocks.push_back(b);
		}
		if (blocks.empty()) return false;
		text_.resize(t_total);
		const size_t T = std::min<size_t>((size_t)threads_, blocks.size());
		std::vector<uint8_t> bad(T, 0);
		auto work = [&](size_t t) {
			z_stream zs;
			memset(&zs, 0, sizeof zs);
			if (inflateInit2(&zs, -15) != Z_OK) { bad[t] = 1; return; }
			for (size_t k = t; k < blocks.size(); k += T) {
				const Block &b = blocks[k];
				zs.next_in = comp_.data() + b.c_off; zs.avail_in = (uInt)b.c_len;
				zs.next_out = (Bytef*)text_.data() + b.t_off; zs.avail_out = (uInt)b.t_len;
				const int rc = inflate(&zs, Z_FINISH);
				if (!(rc == Z_STREAM_END || (rc == Z_OK && b.t_len == 0)) || zs.avail_out != 0 ||
				    (uint32_t)crc32(crc32(0L, Z_NULL, 0), (const Bytef*)text_.data() + b.t_off, (uInt)b.t_len) != b.crc) { bad[t] = 1; break; }
				inflateReset(&zs);
			}
			inflateEnd(&zs);
		};
		if (T <= 1) work(0);
		else {
			std::vector<std::thread> th;
			for (size_t t = 1; t < T; ++t) th.emplace_back(work, t);
			work(0);
			for (std::thread &x : th) x.join();
		}
		for (uint8_t b : bad) if (b) { why_ = "the BGZF input is damaged (a block does not inflate to its checksum)"; text_.clear(); return false; }
		return true;
	}

	FILE *f_ = nullptr; bool own_ = false;
	gzFile gz_ = nullptr;
	bool bgzf_ = false, raw_ = false, raw_gzip_ = false, eof_ = false, zs_open_ = false;
	int threads_ = 1;
	z_stream zs_;
	std::vector<uint8_t> pend_, comp_, cbuf_; size_t pend_at_ = 0;
	std::vector<char> text_; size_t text_at_ = 0;
	std::string why_;
};

} // namespace pansvr
