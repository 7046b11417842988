#include "../pcpx_ranges_standin.hpp"
