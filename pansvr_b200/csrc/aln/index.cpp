#include "index.hpp"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <algorithm>
#include <thread>

namespace pansvr {

namespace {

template <class T>
bool slurp(const std::string &dir, const char *fn, std::vector<T> &out, size_t extra_bytes, std::string &err)
{
	std::string path = dir;
	if (!path.empty() && path.back() != '/') path += '/';
	path += fn;
	FILE *f = fopen(path.c_str(), "rb");
	if (!f) { err = "cannot open " + path; return false; }
	fseek(f, 0, SEEK_END);
	const size_t bytes = (size_t)ftell(f);
	rewind(f);
	const size_t n_el = (bytes + extra_bytes + sizeof(T) - 1) / sizeof(T);
	out.clear();
	out.reserve(n_el);
	// read straight into the vector's storage in 64 MiB pieces (no zero fill of the 2 GiB bucket table first)
	size_t got = 0;
	{
		const size_t full = bytes / sizeof(T);
		std::vector<T> chunk;
		const size_t step = ((size_t)64 << 20) / sizeof(T);
		for (size_t done = 0; done < full;) {
			const size_t k = std::min(step, full - done);
			chunk.resize(k);
			const size_t r = fread(chunk.data(), sizeof(T), k, f);
			out.insert(out.end(), chunk.begin(), chunk.begin() + r);
			got += r * sizeof(T);
			done += k;
			if (r != k) break;
		}
		const size_t tail = bytes - full * sizeof(T);
		if (tail) { T last = 0; got += fread(&last, 1, tail, f); out.push_back(last); }
	}
	out.resize(n_el, 0);
	fclose(f);
	if (got != bytes) { err = "short read on " + path; return false; }
	return true;
}

// The bucket table, compacted: non-empty buckets and a directory over their top 20 bits.
bool load_buckets(const std::string &dir, DebgaIndex &ix, std::string &err)
{
	std::string path = dir;
	if (!path.empty() && path.back() != '/') path += '/';
	path += "unipath_g.hash";
	const int fd = open(path.c_str(), O_RDONLY);
	if (fd < 0) { err = "cannot open " + path; return false; }
	struct stat st;
	const size_t n_bkt = (size_t)1 << 28;                     // 4^14 buckets, n_bkt + 1 starts
	if (fstat(fd, &st) != 0 || (size_t)st.st_size < (n_bkt + 1) * 8) { close(fd); err = "short bucket table " + path; return false; }
	const uint64_t *h = (const uint64_t*)mmap(nullptr, (n_bkt + 1) * 8, PROT_READ, MAP_PRIVATE, fd, 0);
	close(fd);
	if (h == MAP_FAILED) { err = "mmap failed on " + path; return false; }
	madvise((void*)h, (n_bkt + 1) * 8, MADV_SEQUENTIAL);
	const unsigned T = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
	std::vector<std::vector<uint32_t>> keys(T);
	std::vector<std::vector<uint64_t>> starts(T);
	std::vector<uint8_t> bad(T, 0);
	{
		std::vector<std::thread> th;
		const size_t per = (n_bkt + T - 1) / T;
		for (unsigned t = 0; t < T; ++t) th.emplace_back([&, t]() {
			const size_t b = std::min(n_bkt, per * t), e = std::min(n_bkt, per * (t + 1));
			uint64_t prev = b < e ? h[b] : 0;
			for (size_t i = b; i < e; ++i) {
				const uint64_t next = h[i + 1];
				if (next != prev) {
					if (next < prev) { bad[t] = 1; return; }
					keys[t].push_back((uint32_t)i); starts[t].push_back(prev);
				}
				prev = next;
			}
		});
		for (std::thread &x : th) x.join();
	}
	const uint64_t total = h[n_bkt];
	munmap((void*)h, (n_bkt + 1) * 8);
	for (uint8_t b : bad) if (b) { err = "bucket table is not monotone: " + path; return false; }
	ix.bkt_key.clear(); ix.bkt_start.clear();
	for (unsigned t = 0; t < T; ++t) {
		ix.bkt_key.insert(ix.bkt_key.end(), keys[t].begin(), keys[t].end());
		ix.bkt_start.insert(ix.bkt_start.end(), starts[t].begin(), starts[t].end());
	}
	ix.bkt_start.push_back(total);
	ix.bkt_dir.assign(((size_t)1 << 20) + 1, 0);
	size_t j = 0;
	for (size_t x = 0; x <= ((size_t)1 << 20); ++x) {
		while (j < ix.bkt_key.size() && (ix.bkt_key[j] >> 8) < x) ++j;
		ix.bkt_dir[x] = (uint32_t)j;
	}
	return true;
}

// strtok("_") semantics of the reference's name parser: empty fields are skipped
std::vector<std::string> split_us(const std::string &s)
{
	std::vector<std::string> out;
	size_t i = 0;
	while (i < s.size()) {
		while (i < s.size() && s[i] == '_') ++i;
		size_t j = i;
		while (j < s.size() && s[j] != '_') ++j;
		if (j > i) out.push_back(s.substr(i, j - i));
		i = j;
	}
	return out;
}

} // namespace

int DebgaIndex::name2id(const std::string &name) const
{
	for (size_t i = 0; i < target_names.size(); ++i)
		if (target_names[i] == name) return (int)i;
	return -1;
}

bool DebgaIndex::load(const std::string &index_dir, const std::string &header_sam, std::string &err)
{
	// the original header: kept verbatim (the reference re-emits header->text), @SQ order defines the ids
	{
		FILE *f = fopen(header_sam.c_str(), "rb");
		if (!f) { err = "cannot open " + header_sam; return false; }
		char buf[1 << 16];
		size_t n;
		header_text.clear();
		while ((n = fread(buf, 1, sizeof buf, f)) > 0) header_text.append(buf, n);
		fclose(f);
		size_t p = 0;
		std::string kept;
		while (p < header_text.size()) {
			size_t e = header_text.find('\n', p);
			if (e == std::string::npos) e = header_text.size();
			const std::string line = header_text.substr(p, e - p);
			if (!line.empty() && line[0] == '@') {
				kept += line; kept += '\n';
				if (line.compare(0, 3, "@SQ") == 0) {
					size_t s = line.find("\tSN:");
					if (s != std::string::npos) {
						size_t t = line.find('\t', s + 4);
						target_names.push_back(line.substr(s + 4, t == std::string::npos ? std::string::npos : t - s - 4));
					}
				}
			}
			p = e + 1;
		}
		header_text = kept;                         // alignment lines of a full SAM are not part of the header
	}
	if (!slurp(index_dir, "ref.seq", ref_seq, 536, err)) return false;          // +536 zero bytes, deBGA_index.cpp:37
	if (!slurp(index_dir, "unipath.seqb", seqb, 8, err)) return false;
	if (!slurp(index_dir, "unipath.seqfb", seqf, 0, err)) return false;
	if (!slurp(index_dir, "unipath.pos", pos, 0, err)) return false;
	if (!slurp(index_dir, "unipath.posp", posp, 0, err)) return false;
	if (!load_buckets(index_dir, *this, err)) return false;
	if (!slurp(index_dir, "unipath_g.kmer", kmer_g, 0, err)) return false;
	if (!slurp(index_dir, "unipath_g.offset", off_g, 0, err)) return false;

	// unipath.chr: alternating name / cumulative end+1 tokens (deBGA_index.cpp:54-72)
	std::string path = index_dir;
	if (!path.empty() && path.back() != '/') path += '/';
	path += "unipath.chr";
	FILE *f = fopen(path.c_str(), "r");
	if (!f) { err = "cannot open " + path; return false; }
	char tok[4096];
	chr_file_n = 0;                                 // sic: the calloc'ed object starts at 0
	chr_names.clear(); chr_end_n.clear();
	uint32_t line_n = 0;
	while (fscanf(f, "%4095s", tok) == 1) {
		if ((line_n & 1) == 0) { chr_names.resize(chr_file_n + 1); chr_names[chr_file_n] = tok; }
		else { chr_end_n.resize(chr_file_n + 1); chr_end_n[chr_file_n++] = (uint32_t)strtoul(tok, 0, 10); }
		++line_n;
	}
	fclose(f);
	if (chr_file_n == 0) { err = "empty unipath.chr"; return false; }
	chr_end_n[0] = 1;                               // START_POS_REF + 1 overwrites anchor 0's end (deBGA_index.cpp:70)
	chr_names.resize(chr_file_n + 1); chr_names[chr_file_n] = "*";
	chr_end_n.resize(chr_file_n + 2, 0);            // the reference's array is zero beyond the loaded part
	reference_len = chr_end_n[chr_file_n - 1];

	// 16 Kbp bucket -> anchor id table (deBGA_index.cpp:355-366)
	chr_search_index.assign((reference_len >> 14) + 2, 0);
	uint32_t filled = 0;
	for (int i = 0; i < chr_file_n; ++i) {
		const uint32_t b = chr_end_n[i] / 0x4000;
		while (b >= filled) chr_search_index[filled++] = (uint32_t)i;
	}
	chr_search_index[filled] = (uint32_t)chr_file_n;

	// anchor name -> SV record (deBGA_index.cpp:410-430, deBGA_index.hpp:74-155)
	sv_info.clear();
	for (int i = 0; i < chr_file_n; ++i) {
		const std::vector<std::string> t = split_us(chr_names[i]);
		if (t.size() < 9) { err = "anchor name is not id_chr_st_len_type_bp1_bp2_end_vcfid: " + chr_names[i]; return false; }
		SvInfo s;
		s.id = (uint32_t)atoi(t[0].c_str());
		s.chr_id = (uint32_t)name2id(t[1]);
		s.st_pos = (uint32_t)atoi(t[2].c_str());
		s.region_len = atoi(t[3].c_str());
		s.sv_type = t[4];
		s.bp1 = (uint64_t)(int64_t)atoi(t[5].c_str());
		s.bp2 = (uint64_t)(int64_t)atoi(t[6].c_str());
		s.ed_pos = (uint64_t)(int64_t)atoi(t[7].c_str());
		s.vcf_id = t[8];
		s.end_offset = (int)(s.ed_pos - s.st_pos - (uint64_t)(int64_t)s.region_len);
		char buf[1200];
		snprintf(buf, sizeof buf, "%d_%d_%ld_%d_%s_%s", (int)s.id, (int)s.chr_id, (long)s.st_pos, s.region_len, s.sv_type.c_str(), s.vcf_id.c_str());
		s.vcf_print = buf;
		sv_info.push_back(s);
	}
	return true;
}

int DebgaIndex::chromosome_id(uint32_t position) const
{
	int file_n = 0;
	const int pos_index = (int)(position / 0x4000);
	int low = (int)chr_search_index[pos_index], high = (int)chr_search_index[pos_index + 1];
	const int pos = (int)position + 1;
	while (low <= high) {
		const int mid = (low + high) >> 1;
		const uint32_t e = chr_end_n[mid] - 1u;       // unsigned compare, as in the reference (int vs uint32_t)
		if ((uint32_t)pos < e) high = mid - 1;
		else if ((uint32_t)pos > e) low = mid + 1;
		else return mid;
		file_n = low;
	}
	return file_n;
}

} // namespace pansvr
