// ref_full_wrapper.cu -- TEST INFRASTRUCTURE: the WHOLE reference program as a library.
//
// One translation unit #includes the reference's three source files verbatim, from where they lie (-I /root/reference):
// APD.cu (device code + RunPatchMatch), APD.cpp (host side: I/O, InuputInitialization, CudaSpaceInitialization, fusion) and
// main.cpp (GenerateSampleList, ComputeRoundNum, ProcessProblem, the round x iteration schedule), against the stub OpenCV /
// Boost headers of oracle/ref_stubs.  Its main() is renamed and exported as ref_main(argc, argv): the reference's own
// end-to-end run -- same flags, same files written -- for tests/test_gpu_cli.py and as an honest end-to-end baseline.
// Deviations, both by macro and none by edit: clock64() in InitRandomStates (APD.cu:916) reads a settable seed; images are
// decoded by the stub cv::imread (binary PGM / PPM content, whatever the file extension says).
// Built by `make -C oracle ref_full` into oracle/_ref/libapd_ref_full.so (git-ignored); needs a GPU at run time.
#include "APD.h"

__device__ long long g_ref_full_seed = 0x1234567LL;
#define clock64() (g_ref_full_seed)
#include "APD.cu"
#undef clock64

#include "APD.cpp"

#define main reference_main
#include "main.cpp"
#undef main

extern "C" {

void ref_full_set_seed(long long seed) { cudaMemcpyToSymbol(g_ref_full_seed, &seed, sizeof(seed)); }

// the reference's main(); returns its exit status (error paths of the reference call exit() themselves)
int ref_main(int argc, char **argv) { return reference_main(argc, argv); }

}  // extern "C"
