// resource.h -- forwarder: the public surface is declared in lens_trace/api.h (see there).
#pragma once
#include "lens_trace/api.h"
