// TEST INFRASTRUCTURE (checker only; never imported by the product path).
// CPython module `_mcubes_ref`: binds the reference's own NumpyMarchingCubes C++ sources, compiled where they lie under
// /root/reference/external/NumpyMarchingCubes/marching_cubes/src (marching_cubes.cpp + pywrapper.cpp), so that tests can call the
// unmodified reference routine `marching_cubes(PyArrayObject*, double isovalue, double truncation)` (pywrapper.cpp:9-54) the way the
// reference's Cython stub does (_mcubes.pyx:20-25).  The stub here replaces only the Cython-generated glue (_mcubes.cpp, generated
// for another CPython version); no reference source is copied.  Built by oracle/Makefile into oracle/_ref/ (git-ignored).
#include <Python.h>
#define PY_ARRAY_UNIQUE_SYMBOL mcubes_PyArray_API
#include "numpy/arrayobject.h"
#include <stdexcept>

PyObject* marching_cubes(PyArrayObject* arr, double isovalue, double truncation);   // pywrapper.cpp:9

static PyObject* py_marching_cubes(PyObject*, PyObject* args) {
    PyObject* vol; double iso, trunc;
    if (!PyArg_ParseTuple(args, "Odd", &vol, &iso, &trunc)) return nullptr;
    if (!PyArray_Check(vol)) { PyErr_SetString(PyExc_TypeError, "volume must be a numpy array"); return nullptr; }
    try {
        return marching_cubes(reinterpret_cast<PyArrayObject*>(vol), iso, trunc);   // (flat vertices f64, flat polygons u64)
    } catch (const std::exception& e) {
        PyErr_SetString(PyExc_RuntimeError, e.what());
        return nullptr;
    }
}

static PyMethodDef methods[] = {{"marching_cubes", py_marching_cubes, METH_VARARGS, "reference NumpyMarchingCubes"}, {nullptr, nullptr, 0, nullptr}};
static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_mcubes_ref", nullptr, -1, methods};
PyMODINIT_FUNC PyInit__mcubes_ref(void) {
    import_array();
    return PyModule_Create(&moddef);
}
