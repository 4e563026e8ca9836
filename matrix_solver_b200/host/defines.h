// defines.h — forwarder: the whole drop-in surface lives in ms_reference_api.h
#pragma once
#include "ms_reference_api.h"
