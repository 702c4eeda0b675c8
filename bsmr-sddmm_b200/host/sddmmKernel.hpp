// Kept for callers of the first round's file name; the reference's header is sddmmKernel.cuh.
#pragma once
#include "sddmmKernel.cuh"
