// Rotating __constant__ banks for the per-network fp32 constants (TcSmall) of the tensor-core MLP kernels.
//
// A __constant__ symbol exists once per device and module, so staging a network's constants into it before a launch
// (cudaMemcpyToSymbolAsync, stream ordered) is only safe while every MLP call of the device goes through ONE stream: a second
// stream's staging copy could overwrite the bank under a kernel that still reads it.  Each symbol therefore holds kConstBanks
// copies and a process-wide table (one per symbol and device, shared by all handles of the device) binds a bank to the stream
// that last used it:
//   * a launch on a stream that already owns a bank reuses it (stream order protects it: the single-stream case costs nothing);
//   * a new stream takes a never-used bank, else steals the least recently used one after making itself wait for everything the
//     previous owner has queued so far (event record on the old stream + cudaStreamWaitEvent; if the old stream no longer
//     exists, a device synchronise).
// Up to kConstBanks streams (or handles on distinct streams) per device therefore run MLP kernels concurrently with no
// ordering imposed between them; beyond that the calls stay correct and serialise on the stolen bank.
#pragma once
#include <cuda_runtime.h>
#include <mutex>

constexpr int kConstBanks = 4;
constexpr int kConstBankDevices = 64;

struct NbConstBankTable {
  std::mutex mu;
  struct Dev {
    bool used[kConstBanks];
    cudaStream_t stream[kConstBanks];
    cudaEvent_t ev[kConstBanks];
    unsigned long long tick[kConstBanks];
    unsigned long long clock;
  } dev[kConstBankDevices];
};

// Bank to stage into for a launch on `st` of `device` (the current device).  cudaSuccess unless the hand-over itself failed.
static inline cudaError_t nb_const_bank_acquire(NbConstBankTable& T, int device, cudaStream_t st, int* bank) {
  std::lock_guard<std::mutex> lock(T.mu);
  NbConstBankTable::Dev& D = T.dev[device % kConstBankDevices];
  ++D.clock;
  for (int b = 0; b < kConstBanks; ++b)
    if (D.used[b] && D.stream[b] == st) { D.tick[b] = D.clock; *bank = b; return cudaSuccess; }
  for (int b = 0; b < kConstBanks; ++b)
    if (!D.used[b]) { D.used[b] = true; D.stream[b] = st; D.tick[b] = D.clock; *bank = b; return cudaSuccess; }
  int v = 0;
  for (int b = 1; b < kConstBanks; ++b) if (D.tick[b] < D.tick[v]) v = b;
  if (!D.ev[v]) { const cudaError_t e = cudaEventCreateWithFlags(&D.ev[v], cudaEventDisableTiming); if (e != cudaSuccess) return e; }
  if (cudaEventRecord(D.ev[v], D.stream[v]) == cudaSuccess) {
    const cudaError_t e = cudaStreamWaitEvent(st, D.ev[v], 0);
    if (e != cudaSuccess) return e;
  } else {
    (void)cudaGetLastError();                    // the previous owner was destroyed: nothing of it can still be queued after this
    const cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return e;
  }
  D.stream[v] = st; D.tick[v] = D.clock; *bank = v;
  return cudaSuccess;
}
