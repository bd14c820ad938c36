// Phase trace of the heads forward kernel on B200 (one CTA's globaltimer stamps + the whole launch by CUDA events).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DHD_TRACE=0 -I cilrs-autonomous-driving-carla_b200/csrc \
//        tools/heads_trace.cu -o tools/heads_trace_test && ./tools/heads_trace_test [batch]
#include "heads_run.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <vector>
namespace cilrs { long long g_cilrs_launches = 0; bool pdl_enabled() { return false; } }
using namespace cilrs;

int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 128;
  HeadsCtx c{};
  const long long sizes[HD_NUM_SLOTS] = {128, 128, 128 * 128, 128,
    256 * 640, 256, 256 * 256, 256, 768, 3, 256 * 640, 256, 256 * 256, 256, 768, 3, 256 * 640, 256, 256 * 256, 256, 768, 3,
    256 * 640, 256, 256 * 256, 256, 768, 3, 256 * 512, 256, 256 * 256, 256, 256, 1};
  long long tot = 0;
  for (int i = 0; i < HD_NUM_SLOTS; ++i) { c.off[i] = tot; tot += (sizes[i] + 15) / 16 * 16; }
  std::vector<float> hp(tot);
  for (long long i = 0; i < tot; ++i) hp[i] = 0.02f * (float)((i * 2654435761u) % 1000) / 1000.f - 0.01f;
  float* params; cudaMalloc(&params, tot * 4); cudaMemcpy(params, hp.data(), tot * 4, cudaMemcpyHostToDevice);
  c.params = params;
  float *feat, *speed, *controls, *ps; long long* cmd;
  cudaMalloc(&feat, (size_t)B * 512 * 4); cudaMalloc(&speed, B * 4); cudaMalloc(&controls, B * 12); cudaMalloc(&ps, B * 4); cudaMalloc(&cmd, B * 8);
  std::vector<float> hf((size_t)B * 512, 0.5f), hs(B, 0.3f); std::vector<long long> hc(B);
  for (int i = 0; i < B; ++i) hc[i] = i % 4;
  cudaMemcpy(feat, hf.data(), hf.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(speed, hs.data(), B * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(cmd, hc.data(), B * 8, cudaMemcpyHostToDevice);
  c.feat = feat;
  float** sv[6] = {&c.hs.s1, &c.hs.sfeat, &c.hs.b1, &c.hs.b2, &c.hs.p1, &c.hs.p2};
  const int w[6] = {128, 128, 256, 256, 256, 256};
  for (int i = 0; i < 6; ++i) cudaMalloc(sv[i], (size_t)B * w[i] * 4);
  int* err; cudaMalloc(&err, 64); cudaMemset(err, 0, 64); c.err_flag = err;
  unsigned int* ctr; cudaMalloc(&ctr, 64); cudaMemset(ctr, 0, 64); c.loss_counter = ctr;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0);
    int st = heads_forward_run(c, B, speed, cmd, controls, ps, 1, 0.f, 1ull, nullptr, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("rep %d status %d kernel %.1f us (grid %d CTAs)\n", rep, st, ms * 1e3, heads_grid(B));
  }
#ifdef HD_TRACE
  unsigned long long t[32];
  cudaMemcpyFromSymbol(t, g_hd_trace, sizeof(t));
  const char* names[12] = {"start", "list", "X+se0", "csync", "se3", "csync", "br0", "csync", "br3", "csync", "br6", "csync"};
  for (int i = 1; i < 12; ++i) printf("  %-6s +%6.2f us\n", names[i], (double)(t[i] - t[i - 1]) * 1e-3);
  printf("  total  %6.2f us\n", (double)(t[11] - t[0]) * 1e-3);
#endif
  float hcn[12]; cudaMemcpy(hcn, controls, 48, cudaMemcpyDeviceToHost);
  printf("controls[0..2] = %g %g %g  (%s)\n", hcn[0], hcn[1], hcn[2], cudaGetErrorString(cudaGetLastError()));
  return 0;
}
