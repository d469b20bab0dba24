// bam_out.hpp -- the BAM side of `fc_aln`'s output (the reference's default; `-S` selects SAM text).
//
// The reference formats every record as SAM text, parses it into a bam1_t with htslib's sam_parse1
// (read_realignment.cpp:479-536) and writes it with sam_write1 -> bam_write1 into a BGZF stream
// (read_realignment.cpp:85-94,165-175; htslib 1.9: sam.c:228-269,519-565,1197-1430; bgzf.c:424-470,
// 1522-1581,1619-1637).  Here the same two steps are functions over the record text the pipeline
// already produces:
//   bam_encode_record   SAM line -> [block_size][bam record]   (what sam_parse1 + bam_write1 emit)
//   BamWriter           header + records -> BGZF file, cutting blocks where bgzf_write / bgzf_flush_try
//                       cut them, raw deflate at zlib's default level, blocks compressed on the
//                       pipeline's helper threads; the file is byte-identical to the reference's.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <functional>
#include <string>
#include <unordered_map>
#include <vector>

namespace pansvr {

struct DebgaIndex;

struct BamHeaderInfo {                            // bam_hdr_t of the `header.sam` argument (sam_hdr_parse, sam.c)
	std::string text;
	std::vector<std::string> names;
	std::vector<uint32_t> lengths;
	std::unordered_map<std::string, int> name2id;
	void parse(const std::string &header_text);
};

// Appends the uncompressed BAM form of one SAM record (no trailing newline) to `out`.
// false = a line sam_parse1 would reject (message in err).
bool bam_encode_record(const char *line, size_t len, const BamHeaderInfo &h, std::vector<uint8_t> &out, std::string &err);

using ParallelFor = std::function<void(size_t, const std::function<void(size_t, size_t, int)>&)>;

class BamWriter {
public:
	BamWriter(FILE *f, ParallelFor par) : f_(f), par_(std::move(par)) {}
	bool write_header(const BamHeaderInfo &h);     // bam_hdr_write: magic, text, targets, then a block flush
	// `recs` = concatenated [block_size][record] items, `sizes[i]` = byte length of item i (4 + block_size)
	bool write_records(const uint8_t *recs, const std::vector<uint32_t> &sizes);
	bool close();                                  // flush + the 28-byte EOF block; does not fclose
private:
	bool flush_blocks(const uint8_t *data, const std::vector<std::pair<size_t, uint32_t>> &blocks);
	FILE *f_;
	ParallelFor par_;
	std::vector<uint8_t> pending_;                 // the open block (bgzf's uncompressed_block up to block_offset)
};

} // namespace pansvr
