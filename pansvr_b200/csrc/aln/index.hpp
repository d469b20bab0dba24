// index.hpp -- query side of the deBGA unipath/k-mer index of the SV anchor reference.
//
// Reads the on-disk format `deBGA index -k 22` writes (deBGA_release/src/index_build.c; file list and
// element widths in SURVEY.md section 3.3) the way the reference's loader does
// (src/PanSVgenerateVCF/deBGA_index.cpp:33-80, 355-430), including the parts of its behaviour that
// are visible in the SAM output:
//   * deBGA_INDEX is xcalloc'ed, so `chr_file_n` starts at 0, not at its in-class initialiser 1
//     (deBGA_index.hpp:176): anchors are stored from slot 0 and chr_end_n[0] is then overwritten with 1
//     (deBGA_index.cpp:70).  Anchors 0 and 1 therefore share id 1 and the origin of anchor 0.
//   * anchor names are split on '_' into the 9 fields fc_anchor_ref writes (deBGA_index.hpp:75-93).
#pragma once
#include <stdint.h>
#include <string>
#include <vector>

namespace pansvr {

struct SvInfo {                       // SV_chr_info, deBGA_index.hpp:74-155
	uint32_t id = 0;
	uint32_t chr_id = 0;              // index of the chromosome in the original header (-1 -> 0xffffffff)
	uint64_t st_pos = 0;
	int region_len = 0;
	std::string sv_type;
	uint64_t bp1 = 0, bp2 = 0, ed_pos = 0;
	int end_offset = 0;
	std::string vcf_id, vcf_print;    // "SV:Z:" payload: id_chr_st_len_type_vcfid
};

struct DebgaIndex {
	// the eight arrays, as stored on disk
	std::vector<uint64_t> ref_seq, seqb, seqf, pos, posp, off_g;
	std::vector<uint32_t> kmer_g;
	// unipath_g.hash holds 4^14+1 bucket starts (2 GiB) of which only the non-empty buckets carry information: it is
	// scanned once through mmap and kept as the sorted list of non-empty buckets plus a 2^20-entry directory over the
	// top 20 bits of the 28-bit bucket number.  Lookups return exactly hash[h], hash[h+1].
	std::vector<uint32_t> bkt_dir;    // bkt_dir[x] = first entry of bkt_key with key >= x << 8;  2^20 + 1 entries
	std::vector<uint32_t> bkt_key;    // bucket numbers h with hash[h+1] > hash[h], ascending
	std::vector<uint64_t> bkt_start;  // hash[h] of those buckets, then the total; one entry more than bkt_key
	// anchors
	int chr_file_n = 0;
	std::vector<std::string> chr_names;
	std::vector<uint32_t> chr_end_n;  // cumulative end+1 per anchor; [0] = 1 after the loader's overwrite
	uint64_t reference_len = 0;
	std::vector<uint32_t> chr_search_index;
	std::vector<SvInfo> sv_info;
	// header of the original BAM: written unchanged in front of the output, names index RNAME
	std::string header_text;
	std::vector<std::string> target_names;

	bool load(const std::string &index_dir, const std::string &header_sam, std::string &err);
	int chromosome_id(uint32_t position) const;                       // deBGA_index.cpp:369-396
	uint32_t chr_end_before(int chr_id) const { return chr_id >= 1 ? chr_end_n[chr_id - 1] : 0u; }   // chr_end_n[chr_ID-1], RR:291
	void refseq(uint8_t *out, uint32_t len, uint32_t start) const     // deBGA_index.cpp:307-315
	{
		for (uint32_t m = 0; m < len; ++m) out[m] = (ref_seq[(m + start) >> 5] >> ((31 - ((m + start) & 0x1f)) << 1)) & 0x3;
	}
	int name2id(const std::string &name) const;
};

} // namespace pansvr
