#pragma once
#include "jitify/jitify.hpp"
