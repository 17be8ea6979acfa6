#pragma once
#include "ros_stub/core.h"
