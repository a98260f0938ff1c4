// ORACLE -- test infrastructure, NOT product code.  Unit-level KAT driver of the CPU restatement
// (sampler, barycentric): same command line and file protocol as oracle/ref_kat.cpp.
#include "softgl_oracle.h"

int main(int argc, char **argv) { return SoftGL::oracleKatMain(argc, argv); }
