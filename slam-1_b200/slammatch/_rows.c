/*
 * _rows.c -- builds the result of DescriptorMatcher.knnMatch: tuple[nq] of tuple[<= k] of cv2.DMatch.
 *
 * The reference consumes exactly this structure (`for m, n in matches: if m.distance < 0.7 * n.distance`,
 * tracking.py:22-30, keypoint.py:44-51, Point3D.py:40-49).  OpenCV's own binding creates the DMatch objects in
 * C++; doing it from Python costs ~1 us per `cv2.DMatch(q, t, img, d)` call (argument parsing over three
 * overloads), i.e. 2 ms for a 1000-query frame -- 50x the GPU search it wraps.  Here the objects are allocated
 * with the type's own tp_alloc and their cv::DMatch payload {int queryIdx, trainIdx, imgIdx; float distance}
 * is written in place.  slammatch/matcher.py verifies that payload layout against a normally constructed
 * object before it ever uses this module, and falls back to the plain constructor otherwise.
 *
 * Host-side glue only: no arithmetic of the matching path lives here.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>

typedef struct {
    int32_t queryIdx, trainIdx, imgIdx;
    float distance;
} dmatch_payload;

/* make_rows(dmatch_type, payload_offset, idx_addr, dist_addr, keep_addr, nq, k, img_addr, loc_addr)
 *   idx / dist : int32[nq][2] (row-major), -1 = missing neighbour (rows are cut at the first missing one)
 *   keep       : uint8[nq] or 0; rows with keep[i] == 0 are empty (crossCheck=True semantics)
 *   img / loc  : int32[nq][2] or 0; (imgIdx, trainIdx) per entry for a multi-image collection */
static PyObject *make_rows(PyObject *self, PyObject *args)
{
    PyObject *type_obj;
    Py_ssize_t payload_offset, nq;
    unsigned long long idx_addr, dist_addr, keep_addr, img_addr, loc_addr;
    int k;
    if (!PyArg_ParseTuple(args, "OnKKKniKK", &type_obj, &payload_offset, &idx_addr, &dist_addr, &keep_addr, &nq, &k,
                          &img_addr, &loc_addr))
        return NULL;
    if (!PyType_Check(type_obj)) {
        PyErr_SetString(PyExc_TypeError, "first argument must be the DMatch type");
        return NULL;
    }
    PyTypeObject *type = (PyTypeObject *)type_obj;
    if (payload_offset < (Py_ssize_t)sizeof(PyObject) ||
        payload_offset + (Py_ssize_t)sizeof(dmatch_payload) > type->tp_basicsize || k < 1 || k > 2 || nq < 0) {
        PyErr_SetString(PyExc_ValueError, "bad payload offset / k / nq");
        return NULL;
    }
    const int32_t *idx = (const int32_t *)(uintptr_t)idx_addr;
    const int32_t *dist = (const int32_t *)(uintptr_t)dist_addr;
    const uint8_t *keep = (const uint8_t *)(uintptr_t)keep_addr;
    const int32_t *img = (const int32_t *)(uintptr_t)img_addr;
    const int32_t *loc = (const int32_t *)(uintptr_t)loc_addr;

    PyObject *rows = PyTuple_New(nq);
    if (!rows) return NULL;
    for (Py_ssize_t i = 0; i < nq; ++i) {
        int n = 0;
        if (!keep || keep[i])
            while (n < k && idx[2 * i + n] >= 0) ++n;
        PyObject *row = PyTuple_New(n);
        if (!row) { Py_DECREF(rows); return NULL; }
        for (int c = 0; c < n; ++c) {
            PyObject *o = type->tp_alloc(type, 0);
            if (!o) { Py_DECREF(row); Py_DECREF(rows); return NULL; }
            dmatch_payload *p = (dmatch_payload *)((char *)o + payload_offset);
            p->queryIdx = (int32_t)i;
            p->trainIdx = loc ? loc[2 * i + c] : idx[2 * i + c];
            p->imgIdx = img ? img[2 * i + c] : 0;
            p->distance = (float)dist[2 * i + c];
            PyTuple_SET_ITEM(row, c, o);
        }
        PyTuple_SET_ITEM(rows, i, row);
    }
    return rows;
}

static PyMethodDef methods[] = {
    {"make_rows", make_rows, METH_VARARGS, "tuple of tuples of DMatch from idx / dist arrays"},
    {NULL, NULL, 0, NULL},
};

static struct PyModuleDef moduledef = {PyModuleDef_HEAD_INIT, "_rows", "knnMatch result rows built in C", -1, methods};

PyMODINIT_FUNC PyInit__rows(void) { return PyModule_Create(&moduledef); }
