// Forwarding header at the reference's include path; the class lives in kicp/facade_core.hpp.
#pragma once
#include "kicp/facade_core.hpp"
