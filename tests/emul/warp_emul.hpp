// warp_emul.hpp -- TEST INFRASTRUCTURE ONLY.
//
// A 32-lane lock-step "warp" made of ucontext fibres, so that the warp programs in
// pansvr_b200/csrc/*.cuh can be stepped on a machine without a GPU (pytest -m "not gpu" uses it to
// check the kernel's logic against the oracle before any GPU time is spent).  Each lane is a
// fibre running the same function; a collective (shfl / ballot / max / sync) parks the lane
// until all 32 have arrived.  Not a performance tool and never part of the product library.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <ucontext.h>
#include <functional>

class WarpEmul {
public:
	static constexpr int N = 32;
	static int lane() { return self().cur_; }

	static uint32_t shfl(uint32_t v, int src)
	{
		WarpEmul &w = self();
		int g = w.arrive(v);
		return w.slot_[g & 1][src & 31];
	}
	static uint32_t ballot(bool p)
	{
		WarpEmul &w = self();
		int g = w.arrive(p ? 1u : 0u);
		uint32_t m = 0;
		for (int i = 0; i < N; ++i) m |= (w.slot_[g & 1][i] & 1u) << i;
		return m;
	}
	static int wmax(int v)
	{
		WarpEmul &w = self();
		int g = w.arrive((uint32_t)v);
		int m = (int)w.slot_[g & 1][0];
		for (int i = 1; i < N; ++i) if ((int)w.slot_[g & 1][i] > m) m = (int)w.slot_[g & 1][i];
		return m;
	}
	static void sync() { self().arrive(0); }

	// run fn(lane) on 32 fibres until all return
	static void run(const std::function<void()> &fn)
	{
		WarpEmul &w = self();
		w.fn_ = &fn;
		w.gen_ = 0; w.count_ = 0;
		for (int i = 0; i < N; ++i) {
			if (!w.stack_[i]) w.stack_[i] = (char*)malloc(STACK);
			getcontext(&w.ctx_[i]);
			w.ctx_[i].uc_stack.ss_sp = w.stack_[i];
			w.ctx_[i].uc_stack.ss_size = STACK;
			w.ctx_[i].uc_link = &w.sched_;
			makecontext(&w.ctx_[i], (void (*)())&WarpEmul::trampoline, 0);
			w.done_[i] = false;
		}
		int left = N;
		while (left > 0) {
			long before = w.progress_;
			for (int i = 0; i < N; ++i) {
				if (w.done_[i]) continue;
				w.cur_ = i;
				swapcontext(&w.sched_, &w.ctx_[i]);
				if (w.done_[i]) { --left; ++w.progress_; }
			}
			if (left > 0 && w.progress_ == before) {
				fprintf(stderr, "WarpEmul: deadlock (%d lanes parked at a collective the others never reach)\n", left);
				abort();
			}
		}
	}

private:
	static constexpr size_t STACK = 256 * 1024;
	static WarpEmul &self() { static thread_local WarpEmul w; return w; }
	static void trampoline()
	{
		WarpEmul &w = self();
		(*w.fn_)();
		w.done_[w.cur_] = true;
	}
	int arrive(uint32_t v)
	{
		int g = gen_;
		slot_[g & 1][cur_] = v;
		++progress_;
		if (++count_ == N) { count_ = 0; ++gen_; }
		while (gen_ == g) {            // park until the last lane has arrived
			int me = cur_;
			swapcontext(&ctx_[me], &sched_);
			cur_ = me;
		}
		return g;
	}
	ucontext_t sched_, ctx_[N];
	char *stack_[N] = {0};
	bool done_[N];
	const std::function<void()> *fn_ = 0;
	uint32_t slot_[2][N];
	int gen_ = 0, count_ = 0, cur_ = 0;
	long progress_ = 0;
};
