// Shadow header (oracle/_ref build only): net::MpiFile (Code/net/MpiFile.h:17-78) over POSIX
// pread/pwrite so that the reference's LocalPropertyOutput / LocalDistributionInput write and read
// real files.  Open is "collective": every emulated rank opens the same path.
#pragma once
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <filesystem>
#include <memory>
#include <span>
#include <string>
#include "Exception.h"
#include "net/IOCommunicator.h"
namespace hemelb::net {
  class MpiFile {
  public:
    MpiFile() = default;
    static MpiFile Open(const MpiCommunicator& comm, const std::filesystem::path& filename, int mode, MPI_Info = MPI_INFO_NULL) {
      int flags = (mode & MPI_MODE_WRONLY) ? O_WRONLY : O_RDONLY;
      if (mode & MPI_MODE_CREATE) flags |= O_CREAT;
      int fd = ::open(filename.c_str(), flags, 0644);
      if (fd < 0) throw Exception() << "cannot open " << filename.string();
      MpiFile f;
      f.fd = std::shared_ptr<int>(new int(fd), [](int* p) { ::close(*p); delete p; });
      return f;
    }
    void Close() { fd.reset(); }
    void SetView(MPI_Offset, MPI_Datatype, MPI_Datatype, const std::string&, MPI_Info = MPI_INFO_NULL) {}
    MPI_Offset GetSize() const { struct stat st; ::fstat(*fd, &st); return st.st_size; }
    template <typename T, std::size_t N> void Read(std::span<T, N> b, MPI_Status* = nullptr) {
      ReadAt(cursor, b);
      cursor += b.size_bytes();
    }
    template <typename T, std::size_t N> void ReadAt(MPI_Offset off, std::span<T, N> b, MPI_Status* = nullptr) {
      if (::pread(*fd, (void*)b.data(), b.size_bytes(), off) != (ssize_t)b.size_bytes()) throw Exception() << "short read";
    }
    template <typename T, std::size_t N> void WriteAt(MPI_Offset off, std::span<T const, N> b, MPI_Status* = nullptr) {
      if (::pwrite(*fd, (const void*)b.data(), b.size_bytes(), off) != (ssize_t)b.size_bytes()) throw Exception() << "short write";
    }
  private:
    std::shared_ptr<int> fd;
    MPI_Offset cursor = 0;
  };
}
