// SHIM (see glm/glm.hpp in this directory): everything the reference uses lives in glm.hpp
#pragma once
#include <glm/glm.hpp>
