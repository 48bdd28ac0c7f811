// Latency microbenchmarks that size the sequencer (single active lane): results in profiles/microbench/README.md
#include <cstdio>
#include <cuda_runtime.h>
struct Big { const int *a; const double *b; int n; double pad[40]; };
__global__ void k(const __grid_constant__ Big p, int *chain, double *out, long long *t) {
  if (threadIdx.x != 0) return;
  long long t0, t1; double x = out[0]; int idx = 0;
  // dependent global loads (L2 hits after the first pass)
  for (int rep = 0; rep < 2; rep++) { t0 = clock64(); idx = 0; for (int i = 0; i < 256; i++) idx = chain[idx]; t1 = clock64(); }
  t[0] = (t1 - t0) / 256; out[1] = idx;
  // dependent __ldg loads on a tiny array (L1 hits)
  t0 = clock64(); idx = 0; for (int i = 0; i < 256; i++) idx = __ldg(&p.a[idx & 15]); t1 = clock64(); t[1] = (t1 - t0) / 256; out[2] = idx;
  // dependent f64 divisions
  t0 = clock64(); for (int i = 0; i < 256; i++) x = __ddiv_rn(x + 3.0, 1.0000001 + x * 1e-9); t1 = clock64(); t[2] = (t1 - t0) / 256; out[3] = x;
  // dependent f64 add
  t0 = clock64(); for (int i = 0; i < 256; i++) x = __dadd_rn(x, 1.5); t1 = clock64(); t[3] = (t1 - t0) / 256; out[4] = x;
  // param-space pointer load through a generic pointer (what `q.s->field` compiles to)
  const Big *gp = &p; long long acc = 0; const Big *volatile holder = gp;
  t0 = clock64(); for (int i = 0; i < 256; i++) { const Big *vp = holder; acc += (long long)vp->a + vp->n + (acc & 1); } t1 = clock64(); t[4] = (t1 - t0) / 256; out[5] = (double)acc;
  // shared memory via generic pointer
  __shared__ Big sb; sb = p; volatile Big *sp = &sb; acc = 0;
  t0 = clock64(); for (int i = 0; i < 256; i++) acc += (long long)sp->a + sp->n; t1 = clock64(); t[5] = (t1 - t0) / 256; out[6] = (double)acc;
  // relaxed.gpu 128-bit load of an L2-resident line (poll cost)
  unsigned long long lo, hi; acc = 0;
  t0 = clock64(); for (int i = 0; i < 64; i++) { asm volatile("{ .reg .b128 q; ld.relaxed.gpu.global.b128 q, [%2]; mov.b128 {%0, %1}, q; }" : "=l"(lo), "=l"(hi) : "l"(chain + ((acc & 1) * 4)) : "memory"); acc += lo; } t1 = clock64(); t[6] = (t1 - t0) / 64; out[7] = (double)acc;
  // __threadfence cost with one pending store
  t0 = clock64(); for (int i = 0; i < 64; i++) { out[8 + (i & 7)] = x; __threadfence(); } t1 = clock64(); t[7] = (t1 - t0) / 64;
}
int main() {
  int n = 1 << 20; int *h = new int[n]; for (int i = 0; i < n; i++) h[i] = (int)(((long long)i * 7919 + 12345) % n);
  int *chain; double *out; long long *t; int *a;
  cudaMalloc(&chain, n * 4); cudaMemcpy(chain, h, n * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&out, 4096); cudaMemset(out, 0, 4096); cudaMalloc(&t, 64); cudaMalloc(&a, 64); cudaMemset(a, 0, 64);
  Big p{}; p.a = a; p.b = out; p.n = 3;
  k<<<1, 32>>>(p, chain, out, t); cudaDeviceSynchronize();
  long long ht[8]; cudaMemcpy(ht, t, 64, cudaMemcpyDeviceToHost);
  const char *nm[] = {"dependent global load (L2)", "dependent __ldg (L1)", "dependent f64 div", "dependent f64 add", "param load via generic ptr (x2)", "smem load via generic ptr (x2)", "ld.relaxed.gpu.b128 (L2)", "store + __threadfence"};
  for (int i = 0; i < 8; i++) printf("%-36s %lld cycles\n", nm[i], ht[i]);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
