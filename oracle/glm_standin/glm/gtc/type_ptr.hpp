#pragma once
#include "../glm.hpp"  // value_ptr lives in the stand-in's glm.hpp
