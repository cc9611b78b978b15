// pipe_microbench.cu -- measures the issue rate of the integer instructions the alignment
// kernels are built from (B200, sm_100a), to put a denominator under the "integer roofline":
// SURVEY.md 8(d) asks for the DPX / integer peak to be measured, MEASURED_PEAKS.json has none.
// Each kernel runs ITER x 8 independent dependency chains per thread of one instruction (or a
// fixed mix), 148*8 CTAs x 256 threads; result = warp-instructions per clock per SM.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o pipe_microbench pipe_microbench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define ITER 4096
#define CHAINS 8

#define KERNEL(name, BODY)                                                          \
__global__ void name(uint32_t *out, uint32_t s0, uint32_t s1, uint32_t s2) {        \
	uint32_t r[CHAINS];                                                             \
	_Pragma("unroll") for (int i = 0; i < CHAINS; ++i) r[i] = threadIdx.x * 77u + i * s0; \
	uint32_t a = s1 + threadIdx.x, b = s2;                                          \
	for (int it = 0; it < ITER; ++it) {                                             \
		_Pragma("unroll") for (int i = 0; i < CHAINS; ++i) { uint32_t &x = r[i]; BODY }  \
	}                                                                               \
	uint32_t acc = 0;                                                               \
	_Pragma("unroll") for (int i = 0; i < CHAINS; ++i) acc ^= r[i];                 \
	if (acc == 0x12345678u) out[0] = acc;                                           \
}

KERNEL(k_lop3,  asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(a), "r"(b));)
KERNEL(k_imad,  asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(a));)
KERNEL(k_imadhi, asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(a));)
KERNEL(k_shf,   asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));)
KERNEL(k_popc,  asm volatile("popc.b32 %0, %0;" : "+r"(x));)
KERNEL(k_prmt,  asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));)
KERNEL(k_viaddmin, x = __viaddmin_u32(x, a, b);)
KERNEL(k_viaddmin16, x = __viaddmin_u16x2(x, a, b);)
KERNEL(k_vimin3, x = __vimin3_u32(x, a, b);)
KERNEL(k_iadd3, asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(a));)
KERNEL(k_shl,   asm volatile("shl.b32 %0, %0, 1;" : "+r"(x));  asm volatile("xor.b32 %0, %0, %1;" : "+r"(x) : "r"(a));)
KERNEL(k_mix_lop_imad, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(a), "r"(b)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(a));)
KERNEL(k_mix_lop2_imad, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(a), "r"(b)); asm volatile("lop3.b32 %0, %0, %1, %2, 0xe8;" : "+r"(x) : "r"(a), "r"(b)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(a));)
KERNEL(k_mix_lop_imadhi, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(a), "r"(b)); asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(a));)
KERNEL(k_mix_lop_add, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(a), "r"(b)); x = x + x;)
KERNEL(k_mix_lop_viadd, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(a), "r"(b)); x = __viaddmin_u32(x, a, b);)
__global__ void k_lds(uint32_t *out, uint32_t s0, uint32_t s1, uint32_t s2) {
	__shared__ uint32_t tab[1024];
	for (int i = threadIdx.x; i < 1024; i += blockDim.x) tab[i] = (i * s0 + s1) & 1023;
	__syncthreads();
	uint32_t r[CHAINS];
	#pragma unroll
	for (int i = 0; i < CHAINS; ++i) r[i] = (threadIdx.x + i * 32) & 1023;
	for (int it = 0; it < ITER; ++it) {
		#pragma unroll
		for (int i = 0; i < CHAINS; ++i) r[i] = tab[r[i]];
	}
	uint32_t acc = 0;
	#pragma unroll
	for (int i = 0; i < CHAINS; ++i) acc ^= r[i];
	if (acc == 0x12345678u) out[0] = acc;
}

template <typename K> static void run(const char *name, K kern, int per_iter, uint32_t *d, int sms, double clk_hz) {
	const int grid = sms * 8, block = 256;
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	kern<<<grid, block>>>(d, 3, 5, 7);
	cudaDeviceSynchronize();
	float best = 1e30f;
	for (int rep = 0; rep < 3; ++rep) {
		cudaEventRecord(e0); kern<<<grid, block>>>(d, 3, 5, 7); cudaEventRecord(e1); cudaEventSynchronize(e1);
		float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
	}
	double winst = (double)grid * (block / 32) * ITER * CHAINS * per_iter;
	double per_s = winst / (best * 1e-3);
	printf("%-22s %8.3f ms  %7.2f Gwarp-inst/s  %6.3f warp-inst/clk/SM (at %.0f MHz)  = %6.1f lanes/clk/SM\n",
		name, best, per_s / 1e9, per_s / sms / clk_hz, clk_hz / 1e6, per_s / sms / clk_hz * 32);
}

int main() {
	cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
	int sms = p.multiProcessorCount;
	int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
	double clk = khz * 1e3;
	printf("%s, %d SMs, nominal max clock %.0f MHz (rates below assume that clock; see clocks in bench.py)\n", p.name, sms, clk / 1e6);
	uint32_t *d; cudaMalloc(&d, 64);
	run("LOP3", k_lop3, 1, d, sms, clk);
	run("IMAD (mad.lo)", k_imad, 1, d, sms, clk);
	run("IMAD.HI (mad.hi)", k_imadhi, 1, d, sms, clk);
	run("SHF (funnel)", k_shf, 1, d, sms, clk);
	run("POPC", k_popc, 1, d, sms, clk);
	run("PRMT", k_prmt, 1, d, sms, clk);
	run("VIADDMNMX.U32", k_viaddmin, 1, d, sms, clk);
	run("VIADDMNMX.U16x2", k_viaddmin16, 1, d, sms, clk);
	run("VIMNMX3.U32", k_vimin3, 1, d, sms, clk);
	run("add.u32 (ptxas picks)", k_iadd3, 1, d, sms, clk);
	run("shl+xor", k_shl, 2, d, sms, clk);
	run("LOP3+IMAD 1:1", k_mix_lop_imad, 2, d, sms, clk);
	run("LOP3+LOP3+IMAD 2:1", k_mix_lop2_imad, 3, d, sms, clk);
	run("LOP3+IMAD.HI 1:1", k_mix_lop_imadhi, 2, d, sms, clk);
	run("LOP3+(x+x) 1:1", k_mix_lop_add, 2, d, sms, clk);
	run("LOP3+VIADDMNMX 1:1", k_mix_lop_viadd, 2, d, sms, clk);
	run("LDS (dependent)", k_lds, 1, d, sms, clk);
	return 0;
}
