// TEST INFRASTRUCTURE ONLY: forwards to the oracle GDAL shim.
#pragma once
#include "gdal.h"
