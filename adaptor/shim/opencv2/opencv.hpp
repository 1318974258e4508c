// see ../opencv2/core/core.hpp
#pragma once
#include "core/core.hpp"
