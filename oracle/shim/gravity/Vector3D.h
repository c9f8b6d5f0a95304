/* Vector3D.h -- see Space.h in this directory (stand-in for N-BodyShop/utility structures/).  TEST INFRASTRUCTURE ONLY. */
#include "Space.h"
