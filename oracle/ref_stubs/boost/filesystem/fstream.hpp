#include "../filesystem.hpp"
