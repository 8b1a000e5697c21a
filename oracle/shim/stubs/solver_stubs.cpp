// TEST INFRASTRUCTURE (oracle/): minimal stand-in for the Eigen headers the reference
// includes (Eigen is not installed in this image and cannot be fetched). It lets the
// UNMODIFIED reference sources under /root/reference compile in place into oracle/_ref/.
// Written for this repo; it is not Eigen code and is never linked into the product.
#include <cstdlib>
#include <cstdio>
#include "MumpsSolver.hpp"
#include "PetscSolver.hpp"
MumpsSolver::MumpsSolver(unsigned int o, bool f):Option(o),Flag(f){ std::fprintf(stderr,"MUMPS unavailable\n"); std::abort(); }
MumpsSolver::~MumpsSolver(){} bool MumpsSolver::SolveSystem(Eigen::SparseMatrix<double>&,Eigen::VectorXd&){return true;} const Eigen::VectorXd& MumpsSolver::GetSolution(){return x;}
PetscSolver::PetscSolver(unsigned int a,unsigned int b,double t,unsigned int):d_nz(a),o_nz(b),Tolerance(t){ std::fprintf(stderr,"PETSc unavailable\n"); std::abort(); }
PetscSolver::~PetscSolver(){} bool PetscSolver::SolveSystem(Eigen::SparseMatrix<double>&,Eigen::VectorXd&){return true;} const Eigen::VectorXd& PetscSolver::GetSolution(){return x;}
