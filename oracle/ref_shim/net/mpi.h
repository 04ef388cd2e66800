// Shadow header (oracle/_ref build only): the few MPI names the reference's extraction sources
// mention (LocalPropertyOutput.cc, LocalDistributionInput.cc); file I/O is done with pread/pwrite
// by the shadow net/MpiFile.h, collectives by the thread-emulated net/IOCommunicator.h.
#pragma once
#include <cstdint>
using MPI_Offset = long long;
using MPI_Info = int;
using MPI_Datatype = int;
using MPI_Op = int;
struct MPI_Status {};
#define MPI_INFO_NULL 0
#define MPI_STATUS_IGNORE nullptr
#define MPI_SUM 1
#define MPI_CHAR 1
#define MPI_MODE_RDONLY 2
#define MPI_MODE_WRONLY 4
#define MPI_MODE_CREATE 1
#define MPI_MODE_EXCL 64
