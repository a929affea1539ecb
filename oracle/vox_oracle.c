#include "vox_oracle.h"
