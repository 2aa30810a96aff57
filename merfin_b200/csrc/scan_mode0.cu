// k_scan<0> instantiation (see scan_kernel.cuh: bit 0 = dump, bit 1 = joint table).
#include "scan_kernel.cuh"

MFN_SCAN_INSTANCE(0)
