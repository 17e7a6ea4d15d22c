// ORACLE - TEST INFRASTRUCTURE ONLY.  The prior edges include this header and use nothing of it.
#pragma once
#include <g2o/types/slam3d/types_slam3d.h>
