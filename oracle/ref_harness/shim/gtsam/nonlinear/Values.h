// ORACLE tier B shim (test infrastructure): forwards to llref_shim.hpp
#pragma once
#include "llref_shim.hpp"
