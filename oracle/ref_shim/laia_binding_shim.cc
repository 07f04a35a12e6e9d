// TEST INFRASTRUCTURE.  pybind11 module exposing the reference's LaiaScheduler exactly as
// laia/src/python_binding.cc:10-14 does.  The reference's own binding file also binds
// TopkScheduler, whose header needs boost::interprocess (absent here), so it cannot be compiled.
#include "laia_scheduler.h"
#include "binding.h"

using namespace laia_cache;

PYBIND11_MODULE(laia_cache, m) {
    py::class_<LaiaScheduler>(m, "LaiaScheduler")
        .def(py::init<>())
        .def("start", &LaiaScheduler::start)
        .def("pop", &LaiaScheduler::pop)
        .def("length", &LaiaScheduler::queue_length);
}
