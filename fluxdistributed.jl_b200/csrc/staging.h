// Native pinned host->device staging ring (reference uploads pageable memory on the loader task,
// src/ddp_tasks.jl:279-282).  See staging.cpp.
#pragma once
#include <pybind11/pybind11.h>
void fd_bind_staging(pybind11::module_& m);
