// ORACLE - TEST INFRASTRUCTURE ONLY.  g2o generates this export-macro header with CMake; nothing in it is needed here.
#pragma once
#define G2O_TYPES_SLAM3D_API
