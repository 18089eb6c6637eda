/* simt_emu.cpp -- TEST INFRASTRUCTURE ONLY; see simt_emu.h */
#include "simt_emu.h"

namespace simt {

Cta* g_cta = nullptr;
unsigned char* g_dsmem = nullptr;

static const size_t STACK_BYTES = 128 * 1024;

#if defined(__x86_64__)
asm(R"(
.text
.globl simt_switch
.type simt_switch,@function
simt_switch:
	pushq %rbp
	pushq %rbx
	pushq %r12
	pushq %r13
	pushq %r14
	pushq %r15
	movq %rsp, (%rdi)
	movq %rsi, %rsp
	popq %r15
	popq %r14
	popq %r13
	popq %r12
	popq %rbx
	popq %rbp
	ret
.size simt_switch,.-simt_switch
)");
#else
#error "simt_emu: only x86-64 hosts are supported"
#endif

static void trampoline()
{
	Cta* c = g_cta;
	c->body();
	Fiber& f = c->f[c->cur];
	f.done = true;
	c->alive--;
	/* a thread that exits no longer takes part in barriers */
	if (c->alive > 0 && c->bar_arrived >= c->alive) {
		c->bar_arrived = 0;
		c->bar_gen++;
	}
	yield();
	abort(); /* never resumed */
}

void yield()
{
	Cta* c = g_cta;
	int from = c->cur;
	void** save = &c->f[from].sp;
	for (int k = 1; k <= c->n; k++) {
		int i = (from + k) % c->n;
		if (!c->f[i].done) {
			if (i == from) return;
			c->cur = i;
			simt_switch(save, c->f[i].sp);
			return;
		}
	}
	/* nobody left: back to the launcher */
	simt_switch(save, c->main_sp);
}

void collective(unsigned mask, uint64_t v)
{
	Cta* c = g_cta;
	Fiber& f = c->f[c->cur];
	Warp& w = c->warps[f.warp];
	if (!((mask >> f.lane) & 1u)) {
		fprintf(stderr, "simt: lane %d not in collective mask %08x\n", f.lane, mask);
		abort();
	}
	w.vals[f.lane] = v;
	w.arrived |= 1u << f.lane;
	f.coll_ready = false;
	f.coll_mask = mask;
	if ((w.arrived & mask) == mask) {
		/* last to arrive: publish a private snapshot to every participant */
		int base = f.warp * 32;
		for (int i = 0; i < 32; i++) {
			if (!((mask >> i) & 1u)) continue;
			Fiber& g = c->f[base + i];
			if (g.coll_mask != mask) {
				fprintf(stderr, "simt: divergent collective masks %08x vs %08x\n", g.coll_mask, mask);
				abort();
			}
			memcpy(g.coll, w.vals, sizeof(w.vals));
			g.coll_ready = true;
		}
		w.arrived &= ~mask;
		return;
	}
	long spins = 0;
	while (!f.coll_ready) {
		yield();
		if (++spins > 100000000L) { fprintf(stderr, "simt: collective deadlock\n"); abort(); }
	}
}

void run_grid(dim3 grid, dim3 block, size_t smem, std::function<void()> body)
{
	Cta cta;
	int n = (int) (block.x * block.y * block.z);
	cta.n = n;
	cta.f.resize(n);
	cta.bdim = block;
	cta.gdim = grid;
	cta.body = body;
	/* fiber stacks are kept between launches (a test may launch a 512-thread CTA tens of
	 * thousands of times); nested launches do not happen */
	static std::vector<char*> stacks;
	while ((int) stacks.size() < n) stacks.push_back((char*) malloc(STACK_BYTES));
	unsigned char* dsmem = (unsigned char*) aligned_alloc(128, ((smem + 127) / 128 + 1) * 128);
	Cta* saved_cta = g_cta;
	unsigned char* saved_ds = g_dsmem;
	g_dsmem = dsmem;
	g_cta = &cta;

	for (unsigned bz = 0; bz < grid.z; bz++)
	for (unsigned by = 0; by < grid.y; by++)
	for (unsigned bx = 0; bx < grid.x; bx++) {
		cta.bid = uint3{bx, by, bz};
		cta.alive = n;
		cta.bar_arrived = 0;
		cta.bar_gen = 0;
		cta.or_flag = 0;
		memset(cta.nb_arrived, 0, sizeof(cta.nb_arrived));
		memset(cta.nb_gen, 0, sizeof(cta.nb_gen));
		memset(cta.warps, 0, sizeof(cta.warps));
		memset(dsmem, 0xcd, smem);
		for (int i = 0; i < n; i++) {
			Fiber& f = cta.f[i];
			f.done = false;
			f.stack = stacks[i];
			f.tid.x = (unsigned) i % block.x;
			f.tid.y = ((unsigned) i / block.x) % block.y;
			f.tid.z = (unsigned) i / (block.x * block.y);
			f.lane = i & 31;
			f.warp = i >> 5;
			f.coll_ready = false;
			f.coll_mask = 0;
			/* initial frame: six callee-saved registers + return address */
			uintptr_t top = ((uintptr_t) (f.stack + STACK_BYTES)) & ~(uintptr_t) 15;
			void** sp = (void**) (top - 8);     /* so that rsp%16==8 at entry */
			*--sp = (void*) &trampoline;
			for (int k = 0; k < 6; k++) *--sp = nullptr;
			f.sp = (void*) sp;
		}
		cta.cur = 0;
		simt_switch(&cta.main_sp, cta.f[0].sp);
		if (cta.alive != 0) { fprintf(stderr, "simt: CTA ended with %d live threads (deadlock?)\n", cta.alive); abort(); }
	}

	g_cta = saved_cta;
	g_dsmem = saved_ds;
	free(dsmem);
}

} /* namespace simt */
