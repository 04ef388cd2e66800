// Test stand-in for net::InterfaceDelegationNet (Code/net/mixins/InterfaceDelegationNet.h) between the
// harness processes of tests/host_lbm_run.cc: the request calls NeighbouringDataManager::ShareNeeds
// makes, delivered at Dispatch() through files in a directory the processes share (no MPI in this
// image).  Every Dispatch is one round: sends are published as <base>.<round>.<from>.<to>, receives
// poll for theirs.  Test infrastructure only.
#pragma once
#include <chrono>
#include <cstdio>
#include <cstring>
#include <span>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>
#include "units.h"
namespace hemelb::net {
  class InterfaceDelegationNet {
  public:
    InterfaceDelegationNet(int rank, int size, std::string base) : rank(rank), size(size), base(std::move(base)) {}
    int Size() const { return size; }
    int Rank() const { return rank; }
    void RequestAllToAllSend(std::vector<int>& v) { for (int p = 0; p < size; ++p) Send(&v[p], sizeof(int), p); }
    void RequestAllToAllReceive(std::vector<int>& v) { for (int p = 0; p < size; ++p) Recv(&v[p], sizeof(int), p); }
    template <class T> void RequestSendV(std::span<const T> data, proc_t to) { Send(data.data(), data.size_bytes(), to); }
    template <class T> void RequestReceiveV(std::span<T> data, proc_t from) { Recv(data.data(), data.size_bytes(), from); }
    void Dispatch() {
      for (auto& s : sends) {
        if (s.peer == rank) continue;
        const std::string name = Name(rank, s.peer, s.seq), tmp = name + ".tmp";
        FILE* fh = fopen(tmp.c_str(), "wb");
        if (!fh || (s.bytes.size() && fwrite(s.bytes.data(), 1, s.bytes.size(), fh) != s.bytes.size()))
          throw std::runtime_error("net stand-in: cannot write " + tmp);
        fclose(fh);
        if (rename(tmp.c_str(), name.c_str())) throw std::runtime_error("net stand-in: cannot publish " + name);
      }
      for (auto& r : recvs) {
        if (r.peer == rank) {  // a message to oneself: the matching send of this round
          for (auto& s : sends)
            if (s.peer == rank && s.seq == r.seq) std::memcpy(r.dst, s.bytes.data(), r.n);
          continue;
        }
        const std::string name = Name(r.peer, rank, r.seq);
        bool got = false;
        for (int tries = 0; tries < 6000 && !got; ++tries) {
          if (FILE* fh = fopen(name.c_str(), "rb")) {
            got = fread(r.dst, 1, r.n, fh) == r.n;
            fclose(fh);
          }
          if (!got) std::this_thread::sleep_for(std::chrono::milliseconds(10));
        }
        if (!got) throw std::runtime_error("net stand-in: nothing arrived in " + name);
      }
      sends.clear();
      recvs.clear();
      ++round;
      std::fill(sendSeq.begin(), sendSeq.end(), 0);
      std::fill(recvSeq.begin(), recvSeq.end(), 0);
    }
  private:
    struct Out { int peer, seq; std::vector<char> bytes; };
    struct In { int peer, seq; void* dst; size_t n; };
    std::string Name(int from, int to, int seq) const {
      return base + ".net." + std::to_string(round) + "." + std::to_string(seq) + "." + std::to_string(from) + "." + std::to_string(to);
    }
    void Send(const void* p, size_t n, int to) {
      if (sendSeq.empty()) { sendSeq.assign(size, 0); recvSeq.assign(size, 0); }
      Out o{to, sendSeq[to]++, std::vector<char>((const char*)p, (const char*)p + n)};
      sends.push_back(std::move(o));
    }
    void Recv(void* p, size_t n, int from) {
      if (sendSeq.empty()) { sendSeq.assign(size, 0); recvSeq.assign(size, 0); }
      recvs.push_back(In{from, recvSeq[from]++, p, n});
    }
    int rank, size, round = 0;
    std::string base;
    std::vector<Out> sends;
    std::vector<In> recvs;
    std::vector<int> sendSeq, recvSeq;
  };
}
