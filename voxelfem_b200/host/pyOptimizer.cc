// pyOptimizer.cc -- pybind11 module `pyOptimizer` (python_bindings/Optimizer.cc:11-23): MMA(numVars, numConstr, xmin, xmax, f, df_dx)
// with setInitialVar / step / enableGCMMA over the device optimizer (vf_mma_*).  f(x) returns the m + 1 values (objective first),
// df_dx(x) the (m + 1) x n array of gradients, as the reference's Eigen ArrayXd / ArrayXXd callbacks do.
#include <pybind11/functional.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>

#include "VoxelFEM.hh"

namespace py = pybind11;
using voxelfem_b200::VXd;
using NpArr = py::array_t<double, py::array::c_style | py::array::forcecast>;
static py::array_t<double> np_of(const VXd &v) { py::array_t<double> a((py::ssize_t)v.size()); std::copy(v.begin(), v.end(), a.mutable_data()); return a; }
static VXd vx_of(const NpArr &a) { return VXd(a.data(), a.data() + a.size()); }

PYBIND11_MODULE(pyOptimizer, m) {
    using voxelfem_b200::MMA;
    py::class_<MMA>(m, "MMA")
        .def(py::init([](int numVars, int numConstr, const NpArr &xmin, const NpArr &xmax, py::function f, py::function df_dx) {
                 auto F = [f](const VXd &x) { py::gil_scoped_acquire gil; return vx_of(NpArr::ensure(f(np_of(x)))); };
                 auto DF = [df_dx](const VXd &x) { py::gil_scoped_acquire gil; return vx_of(NpArr::ensure(df_dx(np_of(x)))); };   // (m + 1) x n, row-major
                 return std::make_unique<MMA>(numVars, numConstr, vx_of(xmin), vx_of(xmax), F, DF);
             }), py::arg("numVars"), py::arg("numConstr"), py::arg("xmin"), py::arg("xmax"), py::arg("f"), py::arg("df_dx"))
        .def("setInitialVar", [](MMA &o, const NpArr &x0) { o.setInitialVar(vx_of(x0)); }, py::arg("x0"))
        .def("step", &MMA::step)
        .def("enableGCMMA", &MMA::enableGCMMA, py::arg("enable"))
        .def("getOptimalVar", [](const MMA &o) { return np_of(o.getOptimalVar()); });
}
