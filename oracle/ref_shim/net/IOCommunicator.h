// Shadow header (oracle/_ref build only): net::IOCommunicator for R emulated ranks, one
// std::thread each.  Only the collectives the extraction sources call (AllReduce/Scan with
// MPI_SUM, Broadcast, Scatter from the IO rank) -- done through a shared scratch array and a
// barrier.  (Reference: Code/net/IOCommunicator.h:22-52, Code/net/MpiCommunicator.h.)
#pragma once
#include <barrier>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>
#include "net/mpi.h"
namespace hemelb::net {
  struct EmulatedWorld {
    explicit EmulatedWorld(int n) : size(n), sync(n), scratch(n), bytes(n) {}
    int size;
    std::barrier<> sync;
    std::vector<std::uint64_t> scratch;
    std::vector<std::vector<char>> bytes;
  };
  class MpiCommunicator {
  public:
    MpiCommunicator(std::shared_ptr<EmulatedWorld> w, int r) : world(std::move(w)), rank(r) {}
    int Rank() const { return rank; }
    int Size() const { return world->size; }
    template <class T> T AllReduce(const T& v, MPI_Op) const {
      world->scratch[rank] = (std::uint64_t)v;
      world->sync.arrive_and_wait();
      T s = 0;
      for (int i = 0; i < world->size; ++i) s += (T)world->scratch[i];
      world->sync.arrive_and_wait();
      return s;
    }
    template <class T> T Scan(const T& v, MPI_Op) const {  // inclusive prefix sum
      world->scratch[rank] = (std::uint64_t)v;
      world->sync.arrive_and_wait();
      T s = 0;
      for (int i = 0; i <= rank; ++i) s += (T)world->scratch[i];
      world->sync.arrive_and_wait();
      return s;
    }
    template <class T> void Broadcast(T& v, int root) const {
      if (rank == root) { world->bytes[0].resize(sizeof(T)); std::memcpy(world->bytes[0].data(), &v, sizeof(T)); }
      world->sync.arrive_and_wait();
      std::memcpy(&v, world->bytes[0].data(), sizeof(T));
      world->sync.arrive_and_wait();
    }
    template <class T> std::vector<T> Scatter(const std::vector<T>& v, int n, int root) const {
      if (rank == root) { world->bytes[0].resize(v.size() * sizeof(T)); std::memcpy(world->bytes[0].data(), v.data(), v.size() * sizeof(T)); }
      world->sync.arrive_and_wait();
      std::vector<T> out(n);
      std::memcpy(out.data(), world->bytes[0].data() + sizeof(T) * n * rank, sizeof(T) * n);
      world->sync.arrive_and_wait();
      return out;
    }
  protected:
    std::shared_ptr<EmulatedWorld> world;
    int rank;
  };
  class IOCommunicator : public MpiCommunicator {
  public:
    using MpiCommunicator::MpiCommunicator;
    static constexpr int IO_RANK = 0;
    bool OnIORank() const { return Rank() == IO_RANK; }
    constexpr int GetIORank() const { return IO_RANK; }
  };
}
