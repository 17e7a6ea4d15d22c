// ORACLE - TEST INFRASTRUCTURE ONLY.
#pragma once
#include <g2o/types/slam3d/types_slam3d.h>
