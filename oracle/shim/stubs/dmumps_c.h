// TEST INFRASTRUCTURE (oracle/): minimal stand-in for the Eigen headers the reference
// includes (Eigen is not installed in this image and cannot be fetched). It lets the
// UNMODIFIED reference sources under /root/reference compile in place into oracle/_ref/.
// Written for this repo; it is not Eigen code and is never linked into the product.
#pragma once
struct DMUMPS_STRUC_C { int job,par,sym,comm_fortran,n,nz_loc; int*irn_loc; int*jcn_loc; double*a_loc; double*rhs; int icntl[60]; int infog[80]; };
inline void dmumps_c(DMUMPS_STRUC_C*){}
