#include "image.h"
