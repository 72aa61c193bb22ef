/* Host-side collate of a list of MolGraphs, as a CPython extension (chemprop/data/collate.py:37-62).
 *
 * The reference concatenates per-molecule numpy arrays in a Python loop; `dmpnn_collate_host` (C ABI) already does the
 * copying in C, but the Python caller still has to collect 4 x n_mols array pointers one by one, which dominates
 * (~12 us per molecule).  This module walks the sequence and the arrays through the CPython / NumPy C API instead, so
 * a batch costs one call per phase:
 *
 *   sizes(mgs)  -> (n_atoms int64[n], n_edges int64[n], d_v, d_e)
 *   fill(mgs, V, E, edge_index, rev, batch [, Vb, Eb, ei32, rev32, batch32])   addresses as ints
 *
 * `fill` writes the public f32 / int64 tensors and, when the five extra addresses are given, the compact transfer copy
 * (bf16 round-to-nearest-even features, int32 indices) in the same pass.  Items must be 4-sequences
 * (V, E, edge_index, rev_edge_index) -- chemprop's `MolGraph` NamedTuple.  Host code only: no GPU work happens here, and
 * `BatchMolGraph` falls back to the ctypes path when this module is not built.
 */
#define PY_SSIZE_T_CLEAN
#define NPY_NO_DEPRECATED_API NPY_1_7_API_VERSION
#include <Python.h>
#include <numpy/arrayobject.h>
#include <stdint.h>
#include <string.h>

static inline uint16_t f32_to_bf16_rne(float f) {
  uint32_t x;
  memcpy(&x, &f, 4);
  if ((x & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((x >> 16) | 0x40u);
  x += 0x7fffu + ((x >> 16) & 1u);
  return (uint16_t)(x >> 16);
}

/* borrowed item i of a fast sequence, checked to be a sequence of >= 4 entries; returns its fast-sequence (new ref) */
static PyObject* mol_fields(PyObject* item) {
  PyObject* f = PySequence_Fast(item, "each molecule must be a (V, E, edge_index, rev_edge_index) sequence");
  if (!f) return NULL;
  if (PySequence_Fast_GET_SIZE(f) < 4) {
    Py_DECREF(f);
    PyErr_SetString(PyExc_ValueError, "each molecule must have the 4 fields V, E, edge_index, rev_edge_index");
    return NULL;
  }
  return f;
}

static PyObject* py_sizes(PyObject* self, PyObject* args) {
  PyObject* mgs;
  if (!PyArg_ParseTuple(args, "O", &mgs)) return NULL;
  PyObject* seq = PySequence_Fast(mgs, "mgs must be a sequence of MolGraphs");
  if (!seq) return NULL;
  const Py_ssize_t n = PySequence_Fast_GET_SIZE(seq);
  npy_intp dims[1] = {n};
  PyArrayObject* na = (PyArrayObject*)PyArray_SimpleNew(1, dims, NPY_INT64);
  PyArrayObject* ne = (PyArrayObject*)PyArray_SimpleNew(1, dims, NPY_INT64);
  if (!na || !ne) { Py_XDECREF(na); Py_XDECREF(ne); Py_DECREF(seq); return NULL; }
  int64_t* pa = (int64_t*)PyArray_DATA(na);
  int64_t* pe = (int64_t*)PyArray_DATA(ne);
  long d_v = 0, d_e = 0;
  for (Py_ssize_t i = 0; i < n; ++i) {
    PyObject* f = mol_fields(PySequence_Fast_GET_ITEM(seq, i));
    if (!f) goto fail;
    PyObject* V = PySequence_Fast_GET_ITEM(f, 0);
    PyObject* E = PySequence_Fast_GET_ITEM(f, 1);
    PyObject* EI = PySequence_Fast_GET_ITEM(f, 2);
    if (!PyArray_Check(V) || !PyArray_Check(E) || !PyArray_Check(EI) || PyArray_NDIM((PyArrayObject*)V) != 2 ||
        PyArray_NDIM((PyArrayObject*)E) != 2 || PyArray_NDIM((PyArrayObject*)EI) != 2) {
      Py_DECREF(f);
      PyErr_SetString(PyExc_ValueError, "V, E, edge_index must be 2-d numpy arrays");
      goto fail;
    }
    pa[i] = (int64_t)PyArray_DIM((PyArrayObject*)V, 0);
    pe[i] = (int64_t)PyArray_DIM((PyArrayObject*)EI, 1);
    if ((int64_t)PyArray_DIM((PyArrayObject*)E, 0) != pe[i]) {
      Py_DECREF(f);
      PyErr_Format(PyExc_ValueError, "MolGraph.E has %lld rows but edge_index has %lld edges",
                   (long long)PyArray_DIM((PyArrayObject*)E, 0), (long long)pe[i]);
      goto fail;
    }
    if (i == 0) {
      d_v = (long)PyArray_DIM((PyArrayObject*)V, 1);
      d_e = (long)PyArray_DIM((PyArrayObject*)E, 1);
    }
    Py_DECREF(f);
  }
  Py_DECREF(seq);
  return Py_BuildValue("(NNll)", (PyObject*)na, (PyObject*)ne, d_v, d_e);
fail:
  Py_DECREF(na);
  Py_DECREF(ne);
  Py_DECREF(seq);
  return NULL;
}

static PyObject* py_fill(PyObject* self, PyObject* args) {
  PyObject* mgs;
  unsigned long long aV, aE, aEI, aRV, aB, cV = 0, cE = 0, cEI = 0, cRV = 0, cB = 0;
  long long d_v, d_e, E_tot;
  if (!PyArg_ParseTuple(args, "OLLLKKKKK|KKKKK", &mgs, &d_v, &d_e, &E_tot, &aV, &aE, &aEI, &aRV, &aB, &cV, &cE, &cEI,
                        &cRV, &cB))
    return NULL;
  PyObject* seq = PySequence_Fast(mgs, "mgs must be a sequence of MolGraphs");
  if (!seq) return NULL;
  float* V_out = (float*)(uintptr_t)aV;
  float* E_out = (float*)(uintptr_t)aE;
  int64_t* ei_out = (int64_t*)(uintptr_t)aEI;
  int64_t* rv_out = (int64_t*)(uintptr_t)aRV;
  int64_t* b_out = (int64_t*)(uintptr_t)aB;
  uint16_t* Vb = (uint16_t*)(uintptr_t)cV;
  uint16_t* Eb = (uint16_t*)(uintptr_t)cE;
  int32_t* ei32 = (int32_t*)(uintptr_t)cEI;
  int32_t* rv32 = (int32_t*)(uintptr_t)cRV;
  int32_t* b32 = (int32_t*)(uintptr_t)cB;
  const int compact = Vb != NULL;
  const Py_ssize_t n = PySequence_Fast_GET_SIZE(seq);
  int64_t a0 = 0, e0 = 0;
  for (Py_ssize_t i = 0; i < n; ++i) {
    PyObject* f = mol_fields(PySequence_Fast_GET_ITEM(seq, i));
    if (!f) { Py_DECREF(seq); return NULL; }
    /* C-contiguous arrays of the right dtype, cast like np.ascontiguousarray(x, dtype) (a no-op INCREF when they already are) */
    PyArrayObject* V = (PyArrayObject*)PyArray_FROM_OTF(PySequence_Fast_GET_ITEM(f, 0), NPY_FLOAT32, NPY_ARRAY_IN_ARRAY | NPY_ARRAY_FORCECAST);
    PyArrayObject* E = (PyArrayObject*)PyArray_FROM_OTF(PySequence_Fast_GET_ITEM(f, 1), NPY_FLOAT32, NPY_ARRAY_IN_ARRAY | NPY_ARRAY_FORCECAST);
    PyArrayObject* EI = (PyArrayObject*)PyArray_FROM_OTF(PySequence_Fast_GET_ITEM(f, 2), NPY_INT64, NPY_ARRAY_IN_ARRAY | NPY_ARRAY_FORCECAST);
    PyArrayObject* RV = (PyArrayObject*)PyArray_FROM_OTF(PySequence_Fast_GET_ITEM(f, 3), NPY_INT64, NPY_ARRAY_IN_ARRAY | NPY_ARRAY_FORCECAST);
    Py_DECREF(f);
    if (!V || !E || !EI || !RV) goto fail_item;
    if (PyArray_NDIM(V) != 2 || PyArray_NDIM(E) != 2 || PyArray_NDIM(EI) != 2 || PyArray_NDIM(RV) != 1 ||
        PyArray_DIM(V, 1) != d_v || PyArray_DIM(E, 1) != d_e || PyArray_DIM(EI, 0) != 2 ||
        PyArray_DIM(E, 0) != PyArray_DIM(EI, 1) || PyArray_DIM(RV, 0) != PyArray_DIM(EI, 1)) {
      PyErr_Format(PyExc_ValueError, "molecule %zd: inconsistent MolGraph shapes", i);
      goto fail_item;
    }
    {
      const int64_t na = (int64_t)PyArray_DIM(V, 0), ne = (int64_t)PyArray_DIM(EI, 1);
      const float* v = (const float*)PyArray_DATA(V);
      const float* e = (const float*)PyArray_DATA(E);
      const int64_t* ei = (const int64_t*)PyArray_DATA(EI);
      const int64_t* rv = (const int64_t*)PyArray_DATA(RV);
      if (na * d_v > 0) memcpy(V_out + a0 * d_v, v, sizeof(float) * (size_t)(na * d_v));
      if (ne * d_e > 0) memcpy(E_out + e0 * d_e, e, sizeof(float) * (size_t)(ne * d_e));
      for (int64_t j = 0; j < ne; ++j) {
        ei_out[e0 + j] = ei[j] + a0;
        ei_out[E_tot + e0 + j] = ei[ne + j] + a0;
        rv_out[e0 + j] = rv[j] + e0;
      }
      for (int64_t j = 0; j < na; ++j) b_out[a0 + j] = (int64_t)i;
      if (compact) {
        for (int64_t j = 0; j < na * d_v; ++j) Vb[a0 * d_v + j] = f32_to_bf16_rne(v[j]);
        for (int64_t j = 0; j < ne * d_e; ++j) Eb[e0 * d_e + j] = f32_to_bf16_rne(e[j]);
        for (int64_t j = 0; j < ne; ++j) {
          ei32[e0 + j] = (int32_t)(ei[j] + a0);
          ei32[E_tot + e0 + j] = (int32_t)(ei[ne + j] + a0);
          rv32[e0 + j] = (int32_t)(rv[j] + e0);
        }
        for (int64_t j = 0; j < na; ++j) b32[a0 + j] = (int32_t)i;
      }
      a0 += na;
      e0 += ne;
    }
    Py_DECREF(V); Py_DECREF(E); Py_DECREF(EI); Py_DECREF(RV);
    continue;
  fail_item:
    Py_XDECREF(V); Py_XDECREF(E); Py_XDECREF(EI); Py_XDECREF(RV);
    Py_DECREF(seq);
    return NULL;
  }
  Py_DECREF(seq);
  if (e0 != E_tot) {
    PyErr_SetString(PyExc_ValueError, "edge count changed between sizes() and fill()");
    return NULL;
  }
  Py_RETURN_NONE;
}

static PyMethodDef methods[] = {
    {"sizes", py_sizes, METH_VARARGS, "per-molecule atom / edge counts and the feature widths"},
    {"fill", py_fill, METH_VARARGS, "concatenate the molecules into preallocated buffers (addresses as ints)"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_collate_ext", "host collate of MolGraph lists", -1, methods};

PyMODINIT_FUNC PyInit__collate_ext(void) {
  import_array();
  return PyModule_Create(&moddef);
}
