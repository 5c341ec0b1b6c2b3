#include "medium.h"
