// ORACLE BUILD SHIM: lower-case spelling used by ClusterLODTypes.h:11.
#pragma once
#include "DirectXMath.h"
