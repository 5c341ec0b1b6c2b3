#include "sampler.h"
