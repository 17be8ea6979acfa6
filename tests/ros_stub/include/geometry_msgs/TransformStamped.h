#pragma once
#include "ros_stub/msgs.h"
