// SHIM: see fmt/core.h in this directory
#pragma once
#include <fmt/core.h>
