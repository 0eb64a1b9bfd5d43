// kernel_extract.cpp -- TEST INFRASTRUCTURE ONLY: prints the reference's OpenCL `traversal` kernel source (the string
// literal traversalKernel in /root/reference/RayAccelerator/Kernels.h:9-242) so that oracle/Makefile can compile it,
// from where it lies, into oracle/_ref/libkernel_ref.so. The text is written to oracle/_ref/ (git-ignored) only.
#include <cstdio>

#include "Kernels.h"

int main() {
	fputs(traversalKernel, stdout);
	return 0;
}
