// Shadow of net/MpiDataType.h (oracle/_ref Domain build only).  The reference maps C++ types to MPI
// datatypes with boost::hana, which this image does not have; the interface is kept (MpiDataType<T>(),
// the Predefined concept, MpiDataTypeRegistrationTraits with its std::array registration) and the
// predefined-type lookup is an overload set instead of a hana map.
#pragma once
#include <array>
#include <cstddef>
#include <cstdint>
#include <type_traits>
#include <mpi.h>
#include "net/MpiError.h"

namespace hemelb::net {
  namespace detail {
    template <typename T> struct predefined { static constexpr bool value = false; };
#define HEMELB_FAKE_BUILTIN(cpp, mpi) \
    template <> struct predefined<cpp> { static constexpr bool value = true; static MPI_Datatype get() { return mpi; } };
    HEMELB_FAKE_BUILTIN(char, MPI_CHAR)
    HEMELB_FAKE_BUILTIN(short, MPI_SHORT)
    HEMELB_FAKE_BUILTIN(int, MPI_INT)
    HEMELB_FAKE_BUILTIN(long, MPI_LONG)
    HEMELB_FAKE_BUILTIN(long long, MPI_LONG_LONG)
    HEMELB_FAKE_BUILTIN(signed char, MPI_SIGNED_CHAR)
    HEMELB_FAKE_BUILTIN(unsigned char, MPI_UNSIGNED_CHAR)
    HEMELB_FAKE_BUILTIN(unsigned short, MPI_UNSIGNED_SHORT)
    HEMELB_FAKE_BUILTIN(unsigned, MPI_UNSIGNED)
    HEMELB_FAKE_BUILTIN(unsigned long, MPI_UNSIGNED_LONG)
    HEMELB_FAKE_BUILTIN(unsigned long long, MPI_UNSIGNED_LONG_LONG)
    HEMELB_FAKE_BUILTIN(float, MPI_FLOAT)
    HEMELB_FAKE_BUILTIN(double, MPI_DOUBLE)
    HEMELB_FAKE_BUILTIN(std::byte, MPI_BYTE)
#undef HEMELB_FAKE_BUILTIN
  }
  template <typename T> concept Predefined = detail::predefined<T>::value;

  template <typename T> struct MpiDataTypeRegistrationTraits;

  template <typename T> MPI_Datatype MpiDataType(T const&) {
    static thread_local MPI_Datatype DT = MPI_DATATYPE_NULL;  // (derived handles are per emulated run)
    DT = MpiDataTypeRegistrationTraits<T>::Register();
    return DT;
  }
  template <Predefined P> MPI_Datatype MpiDataType(P const&) { return detail::predefined<P>::get(); }

  template <typename T> MPI_Datatype MpiDataType() {
    static_assert(std::is_trivial_v<T>, "MPI only works with trivially copyable types and we further require default construction");
    return MpiDataType(T{});
  }

  template <typename T, std::size_t N> struct MpiDataTypeRegistrationTraits<std::array<T, N>> {
    static MPI_Datatype Register() {
      int blocklengths[1] = {N};
      MPI_Datatype types[1] = {MpiDataType<T>()};
      MPI_Aint displacements[1] = {0};
      MPI_Datatype ret;
      MpiCall{MPI_Type_create_struct}(1, blocklengths, displacements, types, &ret);
      MpiCall{MPI_Type_commit}(&ret);
      return ret;
    }
  };
}
