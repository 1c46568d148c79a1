// ofri_comm.cu -- communication back ends of the row-band (domain-decomposed) driver.  Host code only.
//
//   NcclComm   one process per GPU (torchrun): ncclSend/ncclRecv for the ghost rows, ncclAllGather for the coarse flow
//              before the spline up-sample; the scalar reductions (Liu-Shen image maxima and residual sums,
//              Horn-Schunck error sums) go through a one-kernel all-reduce over NVLink peer memory (CUDA IPC mailboxes,
//              below) and fall back to ncclAllReduce when IPC is not available.  libnccl.so.2 is opened with dlopen at first use (the copy
//              already loaded by torch.distributed when there is one), so libofri.so has no link-time NCCL dependency.
//   LocalComm  N bands driven by N host threads of ONE process (on one GPU or on peer-accessible GPUs): stream-ordered
//              device-to-device copies synchronised with CUDA events and host barriers -- no kernel ever waits on
//              another kernel.  This is the back end of the band-invariance tests (N virtual bands on one B200 must
//              reproduce the single-band result bit for bit) and of single-process multi-GPU runs.
// The reference has no counterpart (it is single-process, single-threaded).
#include <dlfcn.h>
#include <stdlib.h>

#include <condition_variable>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "ofri_internal.h"

namespace ofri {

// ---------------------------------------------------------------------------------------------------------------
// Small-payload all-reduce over NVLink peer memory (one process per GPU, CUDA IPC).
//
// The row-band driver all-reduces a handful of scalars many times per solve -- 2 x T residual sums after every fused
// Liu-Shen block (the next block's stopping rule needs them, so the reduction sits on the critical path: 15 per pyramid
// level), the image maxima, the Horn-Schunck error sums.  An NCCL all-reduce of 8 doubles costs a kernel launch plus
// ~20-30 us of protocol at 8 ranks; this one is a single 1-CTA kernel of our own: every rank writes its values straight
// into a mailbox in EVERY peer's HBM (peer stores through the IPC mapping), raises a flag there, waits for the flags of
// all ranks in its own mailbox and sums the contributions in RANK ORDER (so all ranks compute bit-identical results and
// take identical decisions).  Mailboxes are double buffered by the parity of a sequence number: a rank cannot be two
// reductions ahead of a peer (it needs the peer's contribution to the one in between), so parity never collides.
// Ranks run on different GPUs, so the kernels that wait on one another are guaranteed to run concurrently; a generous
// timeout traps instead of hanging the GPU if a peer has died.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kMailRanks = 8;      // one NVSwitch box
constexpr int kMailVals = 64;      // values per reduction
struct PeerMail {
  unsigned long long data[2][kMailRanks][kMailVals];   // [parity][sender][i]: doubles or uint32 as 64-bit words
  unsigned flag[2][kMailRanks];                         // [parity][sender] = sequence number once the data is in place
};
struct PeerMailPtrs { PeerMail* m[kMailRanks]; };

template <bool IS_MAX>
__global__ void __launch_bounds__(kMailVals)
peer_allreduce_kernel(PeerMailPtrs peers, int rank, int nranks, unsigned seq, void* values, int n) {
  const int t = threadIdx.x, par = (int)(seq & 1u);
  unsigned long long mine = 0ull;
  if (t < n) mine = IS_MAX ? (unsigned long long)reinterpret_cast<unsigned*>(values)[t]
                           : reinterpret_cast<unsigned long long*>(values)[t];
  if (t < n)
    for (int r = 0; r < nranks; ++r) peers.m[r]->data[par][rank][t] = mine;      // peer stores over NVLink (r == rank: local)
  __threadfence_system();
  __syncthreads();
  if (t < nranks) {
    *reinterpret_cast<volatile unsigned*>(&peers.m[t]->flag[par][rank]) = seq;  // "rank's data for #seq is in your mailbox"
    volatile unsigned* f = &peers.m[rank]->flag[par][t];                         // wait for rank t's contribution
    long long t0 = clock64();
    while (*f != seq) {
      __nanosleep(200);
      if (clock64() - t0 > 120000000000ll) __trap();                             // ~60 s: a peer died; fail instead of hanging
    }
  }
  __threadfence_system();
  __syncthreads();
  if (t < n) {
    volatile unsigned long long* d = &peers.m[rank]->data[par][0][t];
    if (IS_MAX) {
      unsigned m = 0u;
      for (int r = 0; r < nranks; ++r) {
        const unsigned v = (unsigned)d[(size_t)r * kMailVals];
        m = v > m ? v : m;
      }
      reinterpret_cast<unsigned*>(values)[t] = m;
    } else {
      double acc = 0.0;
      for (int r = 0; r < nranks; ++r) acc += __longlong_as_double((long long)d[(size_t)r * kMailVals]);   // rank order: deterministic
      reinterpret_cast<double*>(values)[t] = acc;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// NCCL through dlopen (minimal declarations of the stable C API, NCCL 2.x)
// ---------------------------------------------------------------------------------------------------------------
namespace {
typedef void* ncclComm_t;
struct ncclUniqueId { char internal[128]; };
enum { kNcclInt8 = 0, kNcclUint32 = 3, kNcclFloat32 = 7, kNcclFloat64 = 8 };
enum { kNcclSum = 0, kNcclMax = 2 };
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  std::string err;
};
NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (!api.lib) {
      api.err = std::string("cannot open libnccl.so.2: ") + (dlerror() ? dlerror() : "?");
      return;
    }
#define OFRI_NCCL_SYM(field, name)                                              \
  *(void**)(&api.field) = dlsym(api.lib, name);                                 \
  if (!api.field) { api.err = std::string("libnccl lacks ") + name; api.lib = nullptr; return; }
    OFRI_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    OFRI_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    OFRI_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    OFRI_NCCL_SYM(Send, "ncclSend")
    OFRI_NCCL_SYM(Recv, "ncclRecv")
    OFRI_NCCL_SYM(AllReduce, "ncclAllReduce")
    OFRI_NCCL_SYM(AllGather, "ncclAllGather")
    OFRI_NCCL_SYM(GroupStart, "ncclGroupStart")
    OFRI_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    OFRI_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef OFRI_NCCL_SYM
  });
  return &api;
}

struct NcclComm : Comm {
  NcclApi* a = nullptr;
  ncclComm_t comm = nullptr;
  std::string err;
  // peer mailboxes of the small-payload all-reduce (null = not available: NCCL does those reductions too)
  PeerMail* mail = nullptr;
  PeerMailPtrs peers = {};
  bool have_mail = false;
  unsigned seq = 0;
  ~NcclComm() override {
    if (have_mail)
      for (int r = 0; r < nranks; ++r)
        if (r != rank && peers.m[r]) cudaIpcCloseMemHandle(peers.m[r]);
    if (mail) cudaFree(mail);
    if (comm) a->CommDestroy(comm);
  }
  // exchange CUDA IPC handles of the mailboxes through the communicator itself; any failure leaves have_mail = false
  void setup_mail() {
    if (nranks < 2 || nranks > kMailRanks) return;
    if (getenv("OFRI_NO_PEER_ALLREDUCE")) return;
    cudaIpcMemHandle_t mine;
    char* d_handles = nullptr;
    std::vector<char> h_handles((size_t)nranks * sizeof(cudaIpcMemHandle_t));
    bool ok = cudaMalloc(&mail, sizeof(PeerMail)) == cudaSuccess && cudaMemset(mail, 0, sizeof(PeerMail)) == cudaSuccess &&
              cudaIpcGetMemHandle(&mine, mail) == cudaSuccess &&
              cudaMalloc(&d_handles, h_handles.size()) == cudaSuccess;
    if (ok)
      ok = cudaMemcpy(d_handles + (size_t)rank * sizeof(mine), &mine, sizeof(mine), cudaMemcpyHostToDevice) == cudaSuccess &&
           a->AllGather(d_handles + (size_t)rank * sizeof(mine), d_handles, sizeof(mine), kNcclInt8, comm, nullptr) == 0 &&
           cudaDeviceSynchronize() == cudaSuccess &&
           cudaMemcpy(h_handles.data(), d_handles, h_handles.size(), cudaMemcpyDeviceToHost) == cudaSuccess;
    // every rank must reach the same verdict: all-reduce "ok" with NCCL itself (min over ranks as max of the negation)
    unsigned* d_flag = nullptr;
    int opened = 0;
    if (ok) {
      for (int r = 0; r < nranks && ok; ++r) {
        if (r == rank) { peers.m[r] = mail; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, h_handles.data() + (size_t)r * sizeof(h), sizeof(h));
        void* p = nullptr;
        ok = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
        peers.m[r] = (PeerMail*)p;
        if (ok) ++opened;
      }
    }
    unsigned bad = ok ? 0u : 1u;
    if (cudaMalloc(&d_flag, sizeof(unsigned)) == cudaSuccess) {
      cudaMemcpy(d_flag, &bad, sizeof(bad), cudaMemcpyHostToDevice);
      if (a->AllReduce(d_flag, d_flag, 1, kNcclUint32, kNcclMax, comm, nullptr) != 0) bad = 1u;
      cudaDeviceSynchronize();
      unsigned any = 1u;
      cudaMemcpy(&any, d_flag, sizeof(any), cudaMemcpyDeviceToHost);
      bad |= any;
      cudaFree(d_flag);
    } else {
      bad = 1u;
    }
    if (d_handles) cudaFree(d_handles);
    cudaGetLastError();
    have_mail = bad == 0u;
    if (!have_mail) {
      for (int r = 0; r < nranks; ++r)
        if (r != rank && peers.m[r]) { cudaIpcCloseMemHandle(peers.m[r]); peers.m[r] = nullptr; }
      (void)opened;
    }
  }
  int check(int rc, const char* what) {
    if (rc == 0) return 0;
    err = std::string(what) + ": " + a->GetErrorString(rc);
    return -1;
  }
  int exchange(int nseg, const float* const* send_up, float* const* recv_up, const float* const* send_dn,
               float* const* recv_dn, size_t count, cudaStream_t s) override {
    if (check(a->GroupStart(), "ncclGroupStart")) return -1;
    int rc = 0;
    for (int i = 0; i < nseg && !rc; ++i) {
      if (rank > 0) {
        rc |= a->Send(send_up[i], count, kNcclFloat32, rank - 1, comm, s);
        rc |= a->Recv(recv_up[i], count, kNcclFloat32, rank - 1, comm, s);
      }
      if (rank < nranks - 1) {
        rc |= a->Send(send_dn[i], count, kNcclFloat32, rank + 1, comm, s);
        rc |= a->Recv(recv_dn[i], count, kNcclFloat32, rank + 1, comm, s);
      }
    }
    int rc2 = a->GroupEnd();
    return check(rc ? rc : rc2, "ncclSend/ncclRecv");
  }
  int allreduce_sum(double* p, size_t n, cudaStream_t s) override {
    if (have_mail && n <= (size_t)kMailVals) {
      peer_allreduce_kernel<false><<<1, kMailVals, 0, s>>>(peers, rank, nranks, ++seq, p, (int)n);
      return cudaPeekAtLastError() == cudaSuccess ? 0 : (err = "peer all-reduce launch failed", -1);
    }
    return check(a->AllReduce(p, p, n, kNcclFloat64, kNcclSum, comm, s), "ncclAllReduce(sum)");
  }
  int allreduce_max_u32(unsigned* p, size_t n, cudaStream_t s) override {
    if (have_mail && n <= (size_t)kMailVals) {
      peer_allreduce_kernel<true><<<1, kMailVals, 0, s>>>(peers, rank, nranks, ++seq, p, (int)n);
      return cudaPeekAtLastError() == cudaSuccess ? 0 : (err = "peer all-reduce launch failed", -1);
    }
    return check(a->AllReduce(p, p, n, kNcclUint32, kNcclMax, comm, s), "ncclAllReduce(max)");
  }
  int allgather(const float* send, float* recv, size_t count, cudaStream_t s) override {
    return check(a->AllGather(send, recv, count, kNcclFloat32, comm, s), "ncclAllGather");
  }
  const char* error() const override { return err.c_str(); }
  bool uses_sms() const override { return true; }
  int peer_allreduce() const override { return have_mail ? 1 : 0; }
};
}  // namespace

int nccl_unique_id(void* out128, std::string* err) {
  NcclApi* a = nccl_api();
  if (!a->lib) { *err = a->err; return -1; }
  ncclUniqueId id;
  int rc = a->GetUniqueId(&id);
  if (rc) { *err = std::string("ncclGetUniqueId: ") + a->GetErrorString(rc); return -1; }
  memcpy(out128, &id, sizeof(id));
  return 0;
}

Comm* make_nccl_comm(int rank, int nranks, const void* uid128, std::string* err) {
  NcclApi* a = nccl_api();
  if (!a->lib) { *err = a->err; return nullptr; }
  ncclUniqueId id;
  memcpy(&id, uid128, sizeof(id));
  NcclComm* c = new NcclComm();
  c->a = a;
  c->rank = rank;
  c->nranks = nranks;
  int rc = a->CommInitRank(&c->comm, nranks, id, rank);
  if (rc) {
    *err = std::string("ncclCommInitRank: ") + a->GetErrorString(rc);
    c->comm = nullptr;
    delete c;
    return nullptr;
  }
  c->setup_mail();      // small-payload all-reduces over peer memory when CUDA IPC works between the ranks
  return c;
}

// ---------------------------------------------------------------------------------------------------------------
// LocalComm: N host threads of one process
// ---------------------------------------------------------------------------------------------------------------
struct LocalGroup {
  int n = 0;
  std::mutex m;
  std::condition_variable cv;
  int count = 0;
  long gen = 0;
  struct Slot {
    const void* p[16];
    cudaEvent_t ready = nullptr, done = nullptr;
    std::vector<double> hd;
    std::vector<unsigned> hu;
  };
  std::vector<Slot> slots;
  bool aborted = false;    // set when a rank failed: wakes every waiter, all later collectives of the group fail
  // false = the group was aborted (a peer failed before or during this collective)
  bool barrier() {
    std::unique_lock<std::mutex> l(m);
    if (aborted) return false;
    long g = gen;
    if (++count == n) {
      count = 0;
      ++gen;
      cv.notify_all();
    } else {
      cv.wait(l, [&] { return gen != g || aborted; });
    }
    return !aborted;
  }
  void abort() {
    { std::lock_guard<std::mutex> l(m); aborted = true; }
    cv.notify_all();
  }
};
LocalGroup* make_local_group(int n) {
  if (n < 1 || n > 64) return nullptr;
  LocalGroup* g = new LocalGroup();
  g->n = n;
  g->slots.resize(n);
  return g;
}
void abort_local_group(LocalGroup* g) {
  if (g) g->abort();
}
void free_local_group(LocalGroup* g) {
  if (!g) return;
  for (auto& s : g->slots) {
    if (s.ready) cudaEventDestroy(s.ready);
    if (s.done) cudaEventDestroy(s.done);
  }
  delete g;
}

namespace {
struct LocalComm : Comm {
  LocalGroup* g = nullptr;
  std::string err;
  int cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    err = std::string(what) + ": " + cudaGetErrorString(e);
    return -1;
  }
  // every rank: record `ready`, meet, pull from the neighbours after waiting for their `ready`, record `done`, meet,
  // wait for the readers' `done` before the send buffers may change, meet (so nobody re-records an event early)
  int exchange(int nseg, const float* const* send_up, float* const* recv_up, const float* const* send_dn,
               float* const* recv_dn, size_t count, cudaStream_t s) override {
    LocalGroup::Slot& me = g->slots[rank];
    if (nseg > 8) { err = "too many segments"; return -1; }
    for (int i = 0; i < nseg; ++i) {
      me.p[i] = send_dn ? send_dn[i] : nullptr;
      me.p[8 + i] = send_up ? send_up[i] : nullptr;
    }
    int rc = cuda(cudaEventRecord(me.ready, s), "cudaEventRecord");
    if (!g->barrier()) { err = "local group aborted: another rank failed"; return -1; }
    const size_t bytes = count * sizeof(float);
    if (!rc && rank > 0) {
      LocalGroup::Slot& nb = g->slots[rank - 1];
      rc |= cuda(cudaStreamWaitEvent(s, nb.ready, 0), "cudaStreamWaitEvent");
      for (int i = 0; i < nseg && !rc; ++i)
        rc |= cuda(cudaMemcpyAsync(recv_up[i], nb.p[i], bytes, cudaMemcpyDefault, s), "cudaMemcpyAsync");
    }
    if (!rc && rank < nranks - 1) {
      LocalGroup::Slot& nb = g->slots[rank + 1];
      rc |= cuda(cudaStreamWaitEvent(s, nb.ready, 0), "cudaStreamWaitEvent");
      for (int i = 0; i < nseg && !rc; ++i)
        rc |= cuda(cudaMemcpyAsync(recv_dn[i], nb.p[8 + i], bytes, cudaMemcpyDefault, s), "cudaMemcpyAsync");
    }
    if (!rc) rc |= cuda(cudaEventRecord(me.done, s), "cudaEventRecord");
    if (!g->barrier()) { err = "local group aborted: another rank failed"; return -1; }
    if (!rc && rank > 0) rc |= cuda(cudaStreamWaitEvent(s, g->slots[rank - 1].done, 0), "cudaStreamWaitEvent");
    if (!rc && rank < nranks - 1) rc |= cuda(cudaStreamWaitEvent(s, g->slots[rank + 1].done, 0), "cudaStreamWaitEvent");
    if (!g->barrier()) { err = "local group aborted: another rank failed"; return -1; }
    return rc;
  }
  template <typename TT, typename Op>
  int allreduce(TT* p, size_t n, cudaStream_t s, std::vector<TT> LocalGroup::Slot::*field, Op op) {
    LocalGroup::Slot& me = g->slots[rank];
    (me.*field).resize(n);
    int rc = cuda(cudaMemcpyAsync((me.*field).data(), p, n * sizeof(TT), cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync");
    rc |= cuda(cudaStreamSynchronize(s), "cudaStreamSynchronize");
    if (!g->barrier()) { err = "local group aborted: another rank failed"; return -1; }
    std::vector<TT> res((g->slots[0].*field).begin(), (g->slots[0].*field).begin() + n);
    for (int r = 1; r < nranks; ++r)
      for (size_t i = 0; i < n; ++i) res[i] = op(res[i], (g->slots[r].*field)[i]);     // rank order: deterministic
    if (!g->barrier()) { err = "local group aborted: another rank failed"; return -1; }
    rc |= cuda(cudaMemcpyAsync(p, res.data(), n * sizeof(TT), cudaMemcpyHostToDevice, s), "cudaMemcpyAsync");
    rc |= cuda(cudaStreamSynchronize(s), "cudaStreamSynchronize");
    return rc;
  }
  int allreduce_sum(double* p, size_t n, cudaStream_t s) override {
    return allreduce<double>(p, n, s, &LocalGroup::Slot::hd, [](double a, double b) { return a + b; });
  }
  int allreduce_max_u32(unsigned* p, size_t n, cudaStream_t s) override {
    return allreduce<unsigned>(p, n, s, &LocalGroup::Slot::hu, [](unsigned a, unsigned b) { return a > b ? a : b; });
  }
  int allgather(const float* send, float* recv, size_t count, cudaStream_t s) override {
    LocalGroup::Slot& me = g->slots[rank];
    me.p[0] = send;
    int rc = cuda(cudaEventRecord(me.ready, s), "cudaEventRecord");
    if (!g->barrier()) { err = "local group aborted: another rank failed"; return -1; }
    for (int r = 0; r < nranks && !rc; ++r) {
      rc |= cuda(cudaStreamWaitEvent(s, g->slots[r].ready, 0), "cudaStreamWaitEvent");
      rc |= cuda(cudaMemcpyAsync(recv + (size_t)r * count, g->slots[r].p[0], count * sizeof(float), cudaMemcpyDefault, s),
                 "cudaMemcpyAsync");
    }
    if (!rc) rc |= cuda(cudaEventRecord(me.done, s), "cudaEventRecord");
    if (!g->barrier()) { err = "local group aborted: another rank failed"; return -1; }
    for (int r = 0; r < nranks && !rc; ++r) rc |= cuda(cudaStreamWaitEvent(s, g->slots[r].done, 0), "cudaStreamWaitEvent");
    if (!g->barrier()) { err = "local group aborted: another rank failed"; return -1; }
    return rc;
  }
  const char* error() const override { return err.c_str(); }
};
}  // namespace

// must be called by the thread that owns rank `rank`, with that rank's device current
Comm* make_local_comm(LocalGroup* g, int rank, std::string* err) {
  if (!g || rank < 0 || rank >= g->n) { *err = "bad local group / rank"; return nullptr; }
  LocalGroup::Slot& s = g->slots[rank];
  if (!s.ready && (cudaEventCreateWithFlags(&s.ready, cudaEventDisableTiming) != cudaSuccess ||
                   cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming) != cudaSuccess)) {
    *err = "cudaEventCreate failed";
    return nullptr;
  }
  LocalComm* c = new LocalComm();
  c->g = g;
  c->rank = rank;
  c->nranks = g->n;
  return c;
}

}  // namespace ofri
