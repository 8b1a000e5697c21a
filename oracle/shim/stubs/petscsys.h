// TEST INFRASTRUCTURE (oracle/): minimal stand-in for the Eigen headers the reference
// includes (Eigen is not installed in this image and cannot be fetched). It lets the
// UNMODIFIED reference sources under /root/reference compile in place into oracle/_ref/.
// Written for this repo; it is not Eigen code and is never linked into the product.
#pragma once
#include "mpi.h"
#define PETSC_COMM_WORLD 0
inline int PetscInitialize(int*,char***,const char*,const char*){return 0;} inline int PetscFinalize(){return 0;}
