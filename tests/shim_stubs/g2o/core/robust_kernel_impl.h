#pragma once
#include "sparse_optimizer.h"
namespace g2o { class RobustKernelHuber : public RobustKernel {}; }
