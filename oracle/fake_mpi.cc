// TEST INFRASTRUCTURE ONLY (oracle/_ref: the reference's geometry::Domain compiled unmodified).
// Thread-backed implementation of oracle/ref_shim_dom/mpi.h: a rank is a thread, a communicator a
// shared object with a barrier, a window is memory that peers copy from and to, a message is a
// heap copy in a mailbox.  Semantics follow the MPI standard for the calls the reference makes on
// this path (net/MpiCommunicator.hpp, net/MpiWindow.h, net/mixins/pointpoint/SeparatedPointPoint.cc,
// net/IOCommunicator.cc); nothing here is taken from the reference or from an MPI implementation.
#include <mpi.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <exception>
#include <map>
#include <mutex>
#include <thread>
#include <tuple>
#include <vector>

namespace {

struct Comm {
  std::vector<int> members;  // world ranks, in rank order of this communicator
  std::mutex m;
  std::condition_variable cv;
  int arrived = 0;
  long generation = 0;
  std::vector<const void*> slots;
  std::vector<long> results;
};
struct Win {
  int comm = 0;
  int dispUnit = 1;
  std::vector<char*> base;
};
struct DerivedType {
  std::vector<std::pair<MPI_Aint, int>> blocks;  // displacement, bytes
  int bytes = 0;
  int kind = 4, elem = 1;
};
struct Message {
  int tag;
  std::vector<char> data;
};
struct Request {
  bool active = false, recv = false;
  void* buf = nullptr;
  int count = 0, src = 0, tag = 0, comm = 0;
  MPI_Datatype type = 0;
};

std::mutex g_lock;  // tables below
std::condition_variable g_mail;
std::deque<Comm*> g_comms;
std::deque<Win*> g_wins;
std::deque<DerivedType> g_types;
std::map<std::tuple<int, int, int>, std::deque<Message>> g_box;  // (comm, source index, destination index)
thread_local int tl_rank = -1;
thread_local std::vector<Request> tl_requests;

[[noreturn]] void die(const char* what) {
  std::fprintf(stderr, "fake MPI: %s\n", what);
  std::abort();
}
Comm& comm_of(MPI_Comm c) {
  std::lock_guard<std::mutex> l(g_lock);
  if (c <= 0 || c >= (int)g_comms.size() || !g_comms[c]) die("invalid communicator");
  return *g_comms[c];
}
int index_in(const Comm& c) {
  for (size_t i = 0; i < c.members.size(); ++i)
    if (c.members[i] == tl_rank) return (int)i;
  die("the calling rank is not in the communicator");
}
void barrier(Comm& c) {
  std::unique_lock<std::mutex> l(c.m);
  const long g = c.generation;
  if (++c.arrived == (int)c.members.size()) {
    c.arrived = 0;
    ++c.generation;
    c.cv.notify_all();
  } else {
    c.cv.wait(l, [&] { return c.generation != g; });
  }
}
// every member's pointer, valid until the closing barrier()
const std::vector<const void*>& publish(Comm& c, const void* mine) {
  c.slots[index_in(c)] = mine;
  barrier(c);
  return c.slots;
}
int new_comm(std::vector<int> members) {
  std::lock_guard<std::mutex> l(g_lock);
  Comm* c = new Comm;
  c->members = std::move(members);
  c->slots.resize(c->members.size());
  c->results.resize(c->members.size());
  g_comms.push_back(c);
  return (int)g_comms.size() - 1;
}

int type_bytes(MPI_Datatype t) {
  if (t >= (1 << 28)) {
    std::lock_guard<std::mutex> l(g_lock);
    return g_types[t - (1 << 28)].bytes;
  }
  return t & 0xffffff;
}
void type_kind(MPI_Datatype t, int& kind, int& elem, int& perItem) {
  if (t >= (1 << 28)) {
    std::lock_guard<std::mutex> l(g_lock);
    const DerivedType& d = g_types[t - (1 << 28)];
    kind = d.kind;
    elem = d.elem;
    perItem = d.bytes / d.elem;
  } else {
    kind = t >> 24;
    elem = t & 0xffffff;
    perItem = 1;
  }
}
// count items of type t at buf -> contiguous bytes
std::vector<char> pack(const void* buf, int count, MPI_Datatype t) {
  std::vector<char> out;
  if (t >= (1 << 28)) {
    DerivedType d;
    {
      std::lock_guard<std::mutex> l(g_lock);
      d = g_types[t - (1 << 28)];
    }
    if (count != 1 && !(d.blocks.size() == 1 && d.blocks[0].first == 0)) die("count > 1 of a struct type");
    for (int k = 0; k < count; ++k)
      for (auto& b : d.blocks) {
        const char* p = (const char*)buf + (MPI_Aint)k * d.bytes + b.first;
        out.insert(out.end(), p, p + b.second);
      }
  } else {
    const char* p = (const char*)buf;
    out.assign(p, p + (size_t)count * type_bytes(t));
  }
  return out;
}
void unpack(const std::vector<char>& in, void* buf, int count, MPI_Datatype t) {
  if (t >= (1 << 28)) {
    DerivedType d;
    {
      std::lock_guard<std::mutex> l(g_lock);
      d = g_types[t - (1 << 28)];
    }
    size_t at = 0;
    for (int k = 0; k < count; ++k)
      for (auto& b : d.blocks) {
        if (at + b.second > in.size()) die("message shorter than the receive type");
        std::memcpy((char*)buf + (MPI_Aint)k * d.bytes + b.first, in.data() + at, b.second);
        at += b.second;
      }
  } else {
    if (in.size() > (size_t)count * type_bytes(t)) die("message longer than the receive buffer");
    std::memcpy(buf, in.data(), in.size());
  }
}

template <class T> void combine_t(T* acc, const T* x, int n, MPI_Op op) {
  for (int i = 0; i < n; ++i) switch (op) {
      case MPI_MAX: acc[i] = std::max(acc[i], x[i]); break;
      case MPI_MIN: acc[i] = std::min(acc[i], x[i]); break;
      case MPI_SUM: acc[i] = acc[i] + x[i]; break;
      case MPI_PROD: acc[i] = acc[i] * x[i]; break;
      case MPI_LOR: acc[i] = (acc[i] || x[i]) ? T(1) : T(0); break;
      case MPI_LAND: acc[i] = (acc[i] && x[i]) ? T(1) : T(0); break;
      default: die("reduction operator not implemented");
    }
}
void combine(void* acc, const void* x, int count, MPI_Datatype t, MPI_Op op) {
  int kind, elem, per;
  type_kind(t, kind, elem, per);
  const int n = count * per;
  if (kind == 1 && elem == 1) combine_t((int8_t*)acc, (const int8_t*)x, n, op);
  else if (kind == 1 && elem == 2) combine_t((int16_t*)acc, (const int16_t*)x, n, op);
  else if (kind == 1 && elem == 4) combine_t((int32_t*)acc, (const int32_t*)x, n, op);
  else if (kind == 1 && elem == 8) combine_t((int64_t*)acc, (const int64_t*)x, n, op);
  else if (kind == 2 && elem == 1) combine_t((uint8_t*)acc, (const uint8_t*)x, n, op);
  else if (kind == 2 && elem == 2) combine_t((uint16_t*)acc, (const uint16_t*)x, n, op);
  else if (kind == 2 && elem == 4) combine_t((uint32_t*)acc, (const uint32_t*)x, n, op);
  else if (kind == 2 && elem == 8) combine_t((uint64_t*)acc, (const uint64_t*)x, n, op);
  else if (kind == 3 && elem == 4) combine_t((float*)acc, (const float*)x, n, op);
  else if (kind == 3 && elem == 8) combine_t((double*)acc, (const double*)x, n, op);
  else die("reduction over an opaque datatype");
}

int unreached(const char* name) {
  std::fprintf(stderr, "fake MPI: %s is not implemented (not on the geometry::Domain path)\n", name);
  std::abort();
}

}  // namespace

extern "C" {

void fakempi_run(int nranks, void (*body)(int, void*), void* arg) {
  {
    std::lock_guard<std::mutex> l(g_lock);
    for (Comm* c : g_comms) delete c;
    for (Win* w : g_wins) delete w;
    g_comms.clear();
    g_wins.clear();
    g_types.clear();
    g_box.clear();
    g_comms.push_back(nullptr);  // MPI_COMM_NULL
    g_wins.push_back(nullptr);   // MPI_WIN_NULL
  }
  std::vector<int> all(nranks);
  for (int r = 0; r < nranks; ++r) all[r] = r;
  if (new_comm(all) != MPI_COMM_WORLD) die("world communicator handle");
  std::vector<std::thread> threads;
  for (int r = 0; r < nranks; ++r)
    threads.emplace_back([=] {
      tl_rank = r;
      tl_requests.assign(1, Request());
      try {
        body(r, arg);
      } catch (std::exception& e) {
        std::fprintf(stderr, "rank %d: exception: %s\n", r, e.what());
        std::abort();
      }
    });
  for (auto& t : threads) t.join();
}

int MPI_Initialized(int* flag) { *flag = 1; return MPI_SUCCESS; }
int MPI_Finalized(int* flag) { *flag = 0; return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm c, int* r) { *r = index_in(comm_of(c)); return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm c, int* n) { *n = (int)comm_of(c).members.size(); return MPI_SUCCESS; }
int MPI_Comm_set_errhandler(MPI_Comm, MPI_Errhandler) { return MPI_SUCCESS; }
int MPI_Comm_free(MPI_Comm* c) { *c = MPI_COMM_NULL; return MPI_SUCCESS; }  // (kept until the next fakempi_run)
int MPI_Comm_compare(MPI_Comm a, MPI_Comm b, int* result) {
  *result = a == b ? MPI_IDENT : (comm_of(a).members == comm_of(b).members ? MPI_CONGRUENT : MPI_UNEQUAL);
  return MPI_SUCCESS;
}
int MPI_Barrier(MPI_Comm c) { barrier(comm_of(c)); return MPI_SUCCESS; }

int MPI_Comm_split(MPI_Comm c, int color, int key, MPI_Comm* out) {
  Comm& C = comm_of(c);
  const int me = index_in(C);
  const int mine[2] = {color, key};
  const auto& all = publish(C, mine);
  std::vector<std::tuple<int, int, int>> group;  // key, old index, world rank
  for (size_t i = 0; i < all.size(); ++i) {
    const int* ck = (const int*)all[i];
    if (ck[0] == color) group.emplace_back(ck[1], (int)i, C.members[i]);
  }
  std::sort(group.begin(), group.end());
  if (color != MPI_UNDEFINED && std::get<1>(group.front()) == me) {
    std::vector<int> members;
    for (auto& g : group) members.push_back(std::get<2>(g));
    const int h = new_comm(members);
    for (auto& g : group) C.results[std::get<1>(g)] = h;
  }
  barrier(C);
  *out = color == MPI_UNDEFINED ? MPI_COMM_NULL : (MPI_Comm)C.results[me];
  barrier(C);
  return MPI_SUCCESS;
}
int MPI_Comm_dup(MPI_Comm c, MPI_Comm* out) { return MPI_Comm_split(c, 0, index_in(comm_of(c)), out); }
// every rank is a thread of one process: one shared-memory node
int MPI_Comm_split_type(MPI_Comm c, int, int key, MPI_Info, MPI_Comm* out) { return MPI_Comm_split(c, 0, key, out); }

int MPI_Bcast(void* buf, int count, MPI_Datatype t, int root, MPI_Comm c) {
  Comm& C = comm_of(c);
  const auto& all = publish(C, buf);
  if (index_in(C) != root) std::memcpy(buf, all[root], (size_t)count * type_bytes(t));
  barrier(C);
  return MPI_SUCCESS;
}
int MPI_Allreduce(const void* send, void* recv, int count, MPI_Datatype t, MPI_Op op, MPI_Comm c) {
  Comm& C = comm_of(c);
  const size_t bytes = (size_t)count * type_bytes(t);
  std::vector<char> mine((const char*)(send == MPI_IN_PLACE ? recv : send), (const char*)(send == MPI_IN_PLACE ? recv : send) + bytes);
  const auto& all = publish(C, mine.data());
  std::memcpy(recv, all[0], bytes);
  for (size_t i = 1; i < all.size(); ++i) combine(recv, all[i], count, t, op);
  barrier(C);
  return MPI_SUCCESS;
}
int MPI_Reduce(const void* send, void* recv, int count, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c) {
  Comm& C = comm_of(c);
  const size_t bytes = (size_t)count * type_bytes(t);
  const bool isRoot = index_in(C) == root;
  std::vector<char> mine((const char*)(send == MPI_IN_PLACE ? recv : send), (const char*)(send == MPI_IN_PLACE ? recv : send) + bytes);
  const auto& all = publish(C, mine.data());
  if (isRoot) {
    std::memcpy(recv, all[0], bytes);
    for (size_t i = 1; i < all.size(); ++i) combine(recv, all[i], count, t, op);
  }
  barrier(C);
  return MPI_SUCCESS;
}
int MPI_Allgatherv(const void* send, int scount, MPI_Datatype st, void* recv, const int* rcounts, const int* displs,
                   MPI_Datatype rt, MPI_Comm c) {
  Comm& C = comm_of(c);
  struct Item { const void* p; size_t bytes; } mine = {send, (size_t)scount * type_bytes(st)};
  const auto& all = publish(C, &mine);
  const int rb = type_bytes(rt);
  for (size_t i = 0; i < all.size(); ++i) {
    const Item* it = (const Item*)all[i];
    if (it->bytes != (size_t)rcounts[i] * rb) die("Allgatherv: counts disagree");
    std::memcpy((char*)recv + (size_t)displs[i] * rb, it->p, it->bytes);
  }
  barrier(C);
  return MPI_SUCCESS;
}
int MPI_Allgather(const void* send, int scount, MPI_Datatype st, void* recv, int rcount, MPI_Datatype rt, MPI_Comm c) {
  const int n = (int)comm_of(c).members.size();
  std::vector<int> counts(n, rcount), displs(n);
  for (int i = 0; i < n; ++i) displs[i] = i * rcount;
  return MPI_Allgatherv(send, scount, st, recv, counts.data(), displs.data(), rt, c);
}
int MPI_Gatherv(const void* send, int scount, MPI_Datatype st, void* recv, const int* rcounts, const int* displs,
                MPI_Datatype rt, int root, MPI_Comm c) {
  Comm& C = comm_of(c);
  struct Item { const void* p; size_t bytes; } mine = {send, (size_t)scount * type_bytes(st)};
  const auto& all = publish(C, &mine);
  if (index_in(C) == root) {
    const int rb = type_bytes(rt);
    for (size_t i = 0; i < all.size(); ++i) {
      const Item* it = (const Item*)all[i];
      std::memcpy((char*)recv + (size_t)displs[i] * rb, it->p, it->bytes);
    }
  }
  barrier(C);
  return MPI_SUCCESS;
}
int MPI_Gather(const void* send, int scount, MPI_Datatype st, void* recv, int rcount, MPI_Datatype rt, int root, MPI_Comm c) {
  const int n = (int)comm_of(c).members.size();
  std::vector<int> counts(n, rcount), displs(n);
  for (int i = 0; i < n; ++i) displs[i] = i * rcount;
  return MPI_Gatherv(send, scount, st, recv, counts.data(), displs.data(), rt, root, c);
}
int MPI_Alltoall(const void* send, int scount, MPI_Datatype st, void* recv, int rcount, MPI_Datatype rt, MPI_Comm c) {
  Comm& C = comm_of(c);
  const int me = index_in(C);
  const auto& all = publish(C, send);
  const size_t sb = (size_t)scount * type_bytes(st);
  if (sb != (size_t)rcount * type_bytes(rt)) die("Alltoall: counts disagree");
  for (size_t i = 0; i < all.size(); ++i) std::memcpy((char*)recv + i * sb, (const char*)all[i] + me * sb, sb);
  barrier(C);
  return MPI_SUCCESS;
}

int MPI_Isend(const void* buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm c, MPI_Request* req) {
  Comm& C = comm_of(c);
  Message m{tag, pack(buf, count, t)};
  {
    std::lock_guard<std::mutex> l(g_lock);
    g_box[{c, index_in(C), dest}].push_back(std::move(m));
  }
  g_mail.notify_all();
  tl_requests.push_back(Request());  // complete at once (eager copy)
  *req = (int)tl_requests.size() - 1;
  return MPI_SUCCESS;
}
int MPI_Irecv(void* buf, int count, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Request* req) {
  Request r;
  r.active = r.recv = true;
  r.buf = buf;
  r.count = count;
  r.type = t;
  r.src = src;
  r.tag = tag;
  r.comm = c;
  tl_requests.push_back(r);
  *req = (int)tl_requests.size() - 1;
  return MPI_SUCCESS;
}
int MPI_Wait(MPI_Request* req, MPI_Status* st) {
  if (*req <= 0) return MPI_SUCCESS;
  Request r = tl_requests[*req];
  if (r.active && r.recv) {
    if (r.src == MPI_ANY_SOURCE) die("receive from any source");
    Comm& C = comm_of(r.comm);
    const int me = index_in(C);
    Message m;
    {
      std::unique_lock<std::mutex> l(g_lock);
      auto& q = g_box[{r.comm, r.src, me}];
      auto match = [&] {
        return std::find_if(q.begin(), q.end(), [&](const Message& x) { return r.tag == MPI_ANY_TAG || x.tag == r.tag; });
      };
      g_mail.wait(l, [&] { return match() != q.end(); });
      auto it = match();
      m = std::move(*it);
      q.erase(it);
    }
    unpack(m.data, r.buf, r.count, r.type);
    if (st) *st = MPI_Status{r.src, m.tag, MPI_SUCCESS};
  }
  tl_requests[*req].active = false;
  *req = MPI_REQUEST_NULL;
  return MPI_SUCCESS;
}
int MPI_Waitall(int n, MPI_Request* reqs, MPI_Status* sts) {
  for (int i = 0; i < n; ++i) MPI_Wait(reqs + i, sts ? sts + i : nullptr);
  return MPI_SUCCESS;
}
int MPI_Send(const void* buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm c) {
  MPI_Request r;
  return MPI_Isend(buf, count, t, dest, tag, c, &r);
}
int MPI_Ssend(const void* buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm c) {
  return MPI_Send(buf, count, t, dest, tag, c);
}
int MPI_Recv(void* buf, int count, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status* st) {
  MPI_Request r;
  MPI_Irecv(buf, count, t, src, tag, c, &r);
  return MPI_Wait(&r, st);
}

int MPI_Type_create_struct(int n, const int* lens, const MPI_Aint* displs, const MPI_Datatype* types, MPI_Datatype* out) {
  DerivedType d;
  for (int i = 0; i < n; ++i) {
    const int b = lens[i] * type_bytes(types[i]);
    d.blocks.emplace_back(displs[i], b);
    d.bytes = std::max<int>(d.bytes, (int)displs[i] + b);
  }
  if (n == 1 && displs[0] == 0) {  // an array of one predefined type keeps its arithmetic
    int per;
    type_kind(types[0], d.kind, d.elem, per);
  }
  std::lock_guard<std::mutex> l(g_lock);
  for (size_t i = 0; i < g_types.size(); ++i)  // (the same definition again: the same handle)
    if (g_types[i].blocks == d.blocks && g_types[i].kind == d.kind && g_types[i].elem == d.elem) {
      *out = (1 << 28) + (int)i;
      return MPI_SUCCESS;
    }
  g_types.push_back(d);
  *out = (1 << 28) + (int)g_types.size() - 1;
  return MPI_SUCCESS;
}
int MPI_Type_commit(MPI_Datatype*) { return MPI_SUCCESS; }
int MPI_Type_free(MPI_Datatype* t) { *t = MPI_DATATYPE_NULL; return MPI_SUCCESS; }
int MPI_Type_size(MPI_Datatype t, int* n) {
  if (t >= (1 << 28)) {
    std::lock_guard<std::mutex> l(g_lock);
    int s = 0;
    for (auto& b : g_types[t - (1 << 28)].blocks) s += b.second;
    *n = s;
  } else {
    *n = type_bytes(t);
  }
  return MPI_SUCCESS;
}
int MPI_Get_address(const void* p, MPI_Aint* a) { *a = (MPI_Aint)p; return MPI_SUCCESS; }
int MPI_Info_create(MPI_Info* i) { *i = 1; return MPI_SUCCESS; }
int MPI_Info_set(MPI_Info, const char*, const char*) { return MPI_SUCCESS; }
int MPI_Info_free(MPI_Info* i) { *i = MPI_INFO_NULL; return MPI_SUCCESS; }

int MPI_Win_allocate(MPI_Aint bytes, int dispUnit, MPI_Info, MPI_Comm c, void* baseptr, MPI_Win* win) {
  Comm& C = comm_of(c);
  const int me = index_in(C);
  char* mem = (char*)std::calloc((size_t)std::max<MPI_Aint>(bytes, 1), 1);
  const auto& all = publish(C, mem);
  if (me == 0) {
    Win* w = new Win;
    w->comm = c;
    w->dispUnit = dispUnit;
    for (auto p : all) w->base.push_back((char*)p);
    std::lock_guard<std::mutex> l(g_lock);
    g_wins.push_back(w);
    for (auto& r : C.results) r = (long)g_wins.size() - 1;
  }
  barrier(C);
  *win = (MPI_Win)C.results[me];
  *(void**)baseptr = mem;
  barrier(C);
  return MPI_SUCCESS;
}
static Win& win_of(MPI_Win w) {
  std::lock_guard<std::mutex> l(g_lock);
  if (w <= 0 || w >= (int)g_wins.size()) die("invalid window");
  return *g_wins[w];
}
int MPI_Win_free(MPI_Win* w) {
  Win& W = win_of(*w);
  Comm& C = comm_of(W.comm);
  barrier(C);  // collective: nobody reads a peer's memory after this
  std::free(W.base[index_in(C)]);
  *w = MPI_WIN_NULL;
  return MPI_SUCCESS;
}
int MPI_Win_fence(int, MPI_Win w) { barrier(comm_of(win_of(w).comm)); return MPI_SUCCESS; }
int MPI_Win_lock(int, int, int, MPI_Win) { return MPI_SUCCESS; }
int MPI_Win_unlock(int, MPI_Win) { return MPI_SUCCESS; }
int MPI_Win_set_errhandler(MPI_Win, MPI_Errhandler) { return MPI_SUCCESS; }
int MPI_Get(void* origin, int ocount, MPI_Datatype ot, int rank, MPI_Aint disp, int, MPI_Datatype, MPI_Win w) {
  Win& W = win_of(w);
  std::memcpy(origin, W.base[rank] + disp * W.dispUnit, (size_t)ocount * type_bytes(ot));
  return MPI_SUCCESS;
}
int MPI_Put(const void* origin, int ocount, MPI_Datatype ot, int rank, MPI_Aint disp, int, MPI_Datatype, MPI_Win w) {
  Win& W = win_of(w);
  std::memcpy(W.base[rank] + disp * W.dispUnit, origin, (size_t)ocount * type_bytes(ot));
  return MPI_SUCCESS;
}
int MPI_Error_string(int, char* s, int* n) { std::strcpy(s, "fake MPI error"); *n = (int)std::strlen(s); return MPI_SUCCESS; }
double MPI_Wtime(void) {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
int MPI_Abort(MPI_Comm, int code) { std::fprintf(stderr, "MPI_Abort(%d)\n", code); std::abort(); }

int MPI_Scan(const void*, void*, int, MPI_Datatype, MPI_Op, MPI_Comm) { return unreached("MPI_Scan"); }
int MPI_Scatter(const void*, int, MPI_Datatype, void*, int, MPI_Datatype, int, MPI_Comm) { return unreached("MPI_Scatter"); }
int MPI_Comm_group(MPI_Comm, MPI_Group*) { return unreached("MPI_Comm_group"); }
int MPI_Comm_create(MPI_Comm, MPI_Group, MPI_Comm*) { return unreached("MPI_Comm_create"); }
int MPI_Group_free(MPI_Group*) { return unreached("MPI_Group_free"); }
int MPI_Group_rank(MPI_Group, int*) { return unreached("MPI_Group_rank"); }
int MPI_Group_size(MPI_Group, int*) { return unreached("MPI_Group_size"); }
int MPI_Group_incl(MPI_Group, int, const int*, MPI_Group*) { return unreached("MPI_Group_incl"); }
int MPI_Group_excl(MPI_Group, int, const int*, MPI_Group*) { return unreached("MPI_Group_excl"); }
int MPI_Group_translate_ranks(MPI_Group, int, const int*, MPI_Group, int*) { return unreached("MPI_Group_translate_ranks"); }
int MPI_Dist_graph_create_adjacent(MPI_Comm, int, const int*, const int*, int, const int*, const int*, MPI_Info, int, MPI_Comm*) { return unreached("MPI_Dist_graph_create_adjacent"); }
int MPI_Dist_graph_neighbors_count(MPI_Comm, int*, int*, int*) { return unreached("MPI_Dist_graph_neighbors_count"); }
int MPI_Dist_graph_neighbors(MPI_Comm, int, int*, int*, int, int*, int*) { return unreached("MPI_Dist_graph_neighbors"); }
int MPI_Neighbor_allgather(const void*, int, MPI_Datatype, void*, int, MPI_Datatype, MPI_Comm) { return unreached("MPI_Neighbor_allgather"); }
int MPI_Neighbor_allgatherv(const void*, int, MPI_Datatype, void*, const int*, const int*, MPI_Datatype, MPI_Comm) { return unreached("MPI_Neighbor_allgatherv"); }
int MPI_Ineighbor_alltoall(const void*, int, MPI_Datatype, void*, int, MPI_Datatype, MPI_Comm, MPI_Request*) { return unreached("MPI_Ineighbor_alltoall"); }
int MPI_Ineighbor_alltoallv(const void*, const int*, const int*, MPI_Datatype, void*, const int*, const int*, MPI_Datatype, MPI_Comm, MPI_Request*) { return unreached("MPI_Ineighbor_alltoallv"); }

}  // extern "C"
