// TEST INFRASTRUCTURE (oracle/): minimal stand-in for the Eigen headers the reference
// includes (Eigen is not installed in this image and cannot be fetched). It lets the
// UNMODIFIED reference sources under /root/reference compile in place into oracle/_ref/.
// Written for this repo; it is not Eigen code and is never linked into the product.
#pragma once
#include <cstring>
typedef int MPI_Comm; typedef int MPI_Datatype; typedef int MPI_Op;
#define MPI_COMM_WORLD 0
#define MPI_DOUBLE 8
#define MPI_SUM 0
inline int MPI_Reduce(const void*s,void*r,int n,MPI_Datatype,MPI_Op,int,MPI_Comm){ std::memcpy(r,s,(size_t)n*8); return 0; }
inline int MPI_Allreduce(const void*s,void*r,int n,MPI_Datatype,MPI_Op,MPI_Comm){ std::memcpy(r,s,(size_t)n*8); return 0; }
inline int MPI_Bcast(void*,int,MPI_Datatype,int,MPI_Comm){ return 0; }
inline int MPI_Barrier(MPI_Comm){ return 0; }
inline int MPI_Comm_rank(MPI_Comm,int*r){*r=0;return 0;} inline int MPI_Comm_size(MPI_Comm,int*s){*s=1;return 0;}
