#include "../core_shim.hpp"
