// stands in for <cuda_runtime.h> when the device sources are compiled for the CPU emulator (tests only)
#pragma once
#include "../cuda_emu.h"
