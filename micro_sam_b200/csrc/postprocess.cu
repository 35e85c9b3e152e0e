#include "engine.h"
