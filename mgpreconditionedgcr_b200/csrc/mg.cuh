// csrc/mg.cuh -- multigrid hierarchy objects behind the opaque mgcr_mg handle (reference: src/MG.h).
#pragma once
#include "ops.cuh"
