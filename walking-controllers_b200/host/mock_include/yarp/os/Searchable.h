// MOCK: see Value.h
#pragma once
#include <yarp/os/Value.h>
