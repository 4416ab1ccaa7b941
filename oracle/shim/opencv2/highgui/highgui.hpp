// oracle shim forwarding header (test infrastructure only)
#pragma once
#include "../cvshim.hpp"
