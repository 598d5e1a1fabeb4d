// Forwarding header at the reference's include path; Config and KinematicICP live in kicp/facade_pipeline.hpp.
#pragma once
#include "kicp/facade_pipeline.hpp"
