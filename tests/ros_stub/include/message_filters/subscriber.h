#pragma once
#include "ros_stub/sync.h"
