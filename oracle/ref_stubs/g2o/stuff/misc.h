// ORACLE - TEST INFRASTRUCTURE ONLY.  g2o/stuff/misc.h: nothing of it is used by plane3d.h.
#pragma once
