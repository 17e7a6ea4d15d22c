#pragma once
#include "edge_se3_prior_stub.h"
