// refrand.hpp -- the two libc random streams the reference's aln stage consumes, reproduced so that
// the replay does not depend on (or disturb) the host process's own rand() state.
//
// The reference never seeds rand(): tie-breaks (read_realignment.cpp:247, read_realignment.hpp:553),
// N-base substitution (read_realignment.cpp:649) and the seeds of the per-handler random_r states
// (read_realignment.hpp:339-340) all draw from glibc's default stream, which is srandom(1) on the
// TYPE_3 additive-feedback generator (degree 31, separation 3).  initstate_r() with a 128-byte buffer
// selects the same TYPE_3 generator.  tests/test_aln_host.py checks both against the real libc.
#pragma once
#include <stdint.h>

namespace pansvr {

class GlibcRandom {
public:
	explicit GlibcRandom(unsigned seed = 1) { reseed(seed); }
	void reseed(unsigned seed)
	{
		if (seed == 0) seed = 1;
		int32_t word = (int32_t)seed;
		r_[0] = word;
		for (int i = 1; i < 31; ++i) {                 // r[i] = 16807 * r[i-1] mod (2^31-1), Schrage's method
			const long hi = word / 127773, lo = word % 127773;
			long w = 16807 * lo - 2836 * hi;
			if (w < 0) w += 2147483647;
			word = (int32_t)w;
			r_[i] = word;
		}
		f_ = 3; b_ = 0;
		for (int i = 0; i < 310; ++i) next();
	}
	int32_t next()                                     // rand() / random_r()
	{
		const uint32_t v = (uint32_t)r_[f_] + (uint32_t)r_[b_];
		r_[f_] = (int32_t)v;
		if (++f_ >= 31) f_ = 0;
		if (++b_ >= 31) b_ = 0;
		return (int32_t)(v >> 1);
	}
	bool operator==(const GlibcRandom &o) const                // same position of the same stream
	{
		for (int i = 0; i < 31; ++i) if (r_[i] != o.r_[i]) return false;
		return f_ == o.f_ && b_ == o.b_;
	}
private:
	int32_t r_[31];
	int f_, b_;
};

} // namespace pansvr
