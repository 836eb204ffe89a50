// Internal declarations shared by the communicator (comm.cpp) and the kernels (allreduce.cu).
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string>

#include "../../include/tok8s.h"

namespace tok {

constexpr int kMaxWorld = TOK_MAX_WORLD;
constexpr int kThreads = 512;   // threads per CTA of every exchange kernel
constexpr int kMaxCtas = 256;   // upper bound on the grid of an exchange kernel (flag slots)

// ---- symmetric heap layout (identical offsets on every replica) -------------------------------
//   [0, kFlagBytes)                      barrier flags   u32 flag[kMaxCtas][kMaxWorld]
//   [64 KiB, ...)                        u64 symoff[kMaxCtas][kMaxWorld]  zero-copy symmetry check
//   [kFlagBytes, +cap)                   staging buffer 0
//   [kFlagBytes + cap, +cap)             staging buffer 1
//   [kFlagBytes + 2*cap, +pool)          symmetric pool: user buckets allocated here (identical
//                                        allocation sequence on every replica => identical offsets)
//                                        are exchanged in place, without staging
constexpr size_t kFlagBytes = 2u << 20;  // one 2 MiB page: keeps staging 2 MiB aligned
constexpr size_t kSymOffBytes = 64u << 10;  // offset of symoff[][] inside the flag page

// local (non-shared) device words, index into KArgs::ctr
constexpr int kCtrCallSeq = kMaxCtas;      // number of completed collective launches
constexpr int kCtrDone = kMaxCtas + 1;     // CTA completion ticket of the running launch
constexpr int kCtrWords = kMaxCtas + 2;

// host-mapped control words (one pinned page), index into KArgs::hostctl
constexpr int kCtlAbort = 0;   // host -> device: leave barriers now
constexpr int kCtlStatus = 1;  // device -> host: 0 ok, 1 timeout, 2 aborted, 3 asymmetric buffer

// internal algorithm ids (zero-copy variants of the public TOK_ALGO_* ones)
constexpr int kAlgoTwoShotInplace = 5;
constexpr int kAlgoNvlsInplace = 6;

struct KArgs {
  const void* in;
  void* out;
  size_t count;            // elements in this launch
  size_t total_packs;      // ceil(count / P)
  size_t packs_per_cta;    // slab length L (multiple of world for two-shot / NVLS)
  size_t stage_off[2];     // byte offsets of the two staging buffers inside a heap
  size_t slot_bytes;       // one-shot: stride between per-source slots inside a staging buffer
  char* peer[kMaxWorld];   // heap base of every rank as mapped in this replica (peer[rank] = own)
  char* mc;                // multicast mapping of the heap, or nullptr
  uint32_t* ctr;           // local device words (kCtrWords)
  volatile uint32_t* hostctl;  // device pointer to the host-mapped control page
  unsigned long long timeout_ns;
  size_t buf_off;              // zero-copy: byte offset of the bucket inside every replica's heap
  unsigned long long* dbg;     // optional [kMaxCtas][8] per-CTA phase timestamps (TOK_DEBUG_PHASES=1)
  float scale;
  int rank;
  int world;
  uint32_t flags;          // TOK_FLAG_SCALE_POST
};

// Implemented in allreduce.cu.  Returns cudaError_t as int (0 = success).
int launch_allreduce(int algo, int in_dtype, int wire_dtype, int out_dtype, int ctas,
                     const KArgs& args, void* stream);
// Elements per 16-byte pack for a dtype triple (4 when any dtype is f32, else 8).
int pack_elems(int in_dtype, int wire_dtype, int out_dtype);
size_t dtype_size(int dtype);

// error plumbing (comm.cpp)
void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);

}  // namespace tok
