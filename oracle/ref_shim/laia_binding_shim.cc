// TEST INFRASTRUCTURE.  pybind11 module `laia_cache` exposing the reference's two planners exactly
// as laia/src/python_binding.cc:8-22 does (same class and method names).  The reference's sources
// (laia/src/laia_scheduler.cc, topk_scheduler.cc, thread_pool.cc, array.cc, utils.cc) are compiled
// unmodified from where they lie; Boost (absent from this image) is replaced by the stand-ins under
// oracle/ref_shim/boost.
#include "laia_scheduler.h"
#include "binding.h"
#include "topk_scheduler.h"

using namespace laia_cache;

PYBIND11_MODULE(laia_cache, m) {
    py::class_<LaiaScheduler>(m, "LaiaScheduler")
        .def(py::init<>())
        .def("start", &LaiaScheduler::start)
        .def("pop", &LaiaScheduler::pop)
        .def("length", &LaiaScheduler::queue_length);

    py::class_<TopkScheduler>(m, "TopkScheduler")
        .def(py::init<>())
        .def("start", &TopkScheduler::start)
        .def("pop", &TopkScheduler::pop)
        .def("pop_from_local_worker", &TopkScheduler::pop_from_local_worker)
        .def("length", &TopkScheduler::queue_length);
}
