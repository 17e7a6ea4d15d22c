// ORACLE - TEST INFRASTRUCTURE ONLY.  CMake-generated export macro of g2o.
#pragma once
#define G2O_TYPES_SLAM3D_ADDONS_API
