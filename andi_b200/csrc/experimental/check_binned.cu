// Compile check of the experimental kernels (make -C andi_b200/csrc experimental). Not linked
// into libandi_b200.so.
#include "walk_binned.cuh"
