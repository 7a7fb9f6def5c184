/* Host-side packing of a poll's reads into the pinned staging arena (riser_b200/preprocess.py RaggedBatch).
 *
 * The ReadUntil client hands over, per channel, the whole accumulated prefix of the current read
 * (riser/client.py:29-31,44-47: AccumulatingCache, np.frombuffer(read.raw_data)).  Gathering 512-3000 of them
 * into one contiguous buffer with a Python loop of slice assignments is the largest part of a live poll's
 * latency (one core's memcpy bandwidth plus ~1 us of interpreter time per read); this module does it with
 * the buffer protocol and a few threads, the GIL released while they copy.
 *
 *   lengths(seq, out_int64)                 bytes of every item -> out[i]
 *   pack(seq, dst, byte_off, skip_bytes, take_bytes, n_threads)
 *       dst[byte_off[i] : byte_off[i] + take_bytes[i]] = item[i][skip_bytes[i] : skip_bytes[i] + take_bytes[i]]
 *
 * Items are any contiguous buffer objects (numpy int16 arrays, bytes).  CPython C API, no numpy headers.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  const char* src;
  char* dst;
  size_t n;
} copy_t;

typedef struct {
  const copy_t* jobs;
  Py_ssize_t lo, hi;
} span_t;

static void* run_span(void* arg) {
  const span_t* s = (const span_t*)arg;
  for (Py_ssize_t i = s->lo; i < s->hi; ++i)
    if (s->jobs[i].n) memcpy(s->jobs[i].dst, s->jobs[i].src, s->jobs[i].n);
  return NULL;
}

static PyObject* hp_lengths(PyObject* self, PyObject* args) {
  PyObject* seq_in;
  Py_buffer out;
  if (!PyArg_ParseTuple(args, "Ow*", &seq_in, &out)) return NULL;
  PyObject* seq = PySequence_Fast(seq_in, "lengths: first argument must be a sequence");
  if (!seq) {
    PyBuffer_Release(&out);
    return NULL;
  }
  const Py_ssize_t n = PySequence_Fast_GET_SIZE(seq);
  if (out.len < (Py_ssize_t)(n * sizeof(int64_t))) {
    PyErr_SetString(PyExc_ValueError, "lengths: output buffer too small");
    goto fail;
  }
  int64_t* o = (int64_t*)out.buf;
  for (Py_ssize_t i = 0; i < n; ++i) {
    Py_buffer v;
    if (PyObject_GetBuffer(PySequence_Fast_GET_ITEM(seq, i), &v, PyBUF_SIMPLE) != 0) goto fail;
    o[i] = (int64_t)v.len;
    PyBuffer_Release(&v);
  }
  Py_DECREF(seq);
  PyBuffer_Release(&out);
  Py_RETURN_NONE;
fail:
  Py_DECREF(seq);
  PyBuffer_Release(&out);
  return NULL;
}

static PyObject* hp_pack(PyObject* self, PyObject* args) {
  PyObject* seq_in;
  Py_buffer dst, off, skip, take;
  int n_threads = 1;
  if (!PyArg_ParseTuple(args, "Ow*y*y*y*|i", &seq_in, &dst, &off, &skip, &take, &n_threads)) return NULL;
  PyObject* seq = PySequence_Fast(seq_in, "pack: first argument must be a sequence");
  Py_buffer* views = NULL;
  copy_t* jobs = NULL;
  Py_ssize_t n = 0, got = 0;
  PyObject* ret = NULL;
  if (!seq) goto done;
  n = PySequence_Fast_GET_SIZE(seq);
  if (off.len < (Py_ssize_t)(n * 8) || skip.len < (Py_ssize_t)(n * 8) || take.len < (Py_ssize_t)(n * 8)) {
    PyErr_SetString(PyExc_ValueError, "pack: offset / skip / take arrays must hold one int64 per item");
    goto done;
  }
  views = (Py_buffer*)calloc((size_t)(n > 0 ? n : 1), sizeof(Py_buffer));
  jobs = (copy_t*)calloc((size_t)(n > 0 ? n : 1), sizeof(copy_t));
  if (!views || !jobs) {
    PyErr_NoMemory();
    goto done;
  }
  const int64_t* o = (const int64_t*)off.buf;
  const int64_t* sk = (const int64_t*)skip.buf;
  const int64_t* tk = (const int64_t*)take.buf;
  size_t total = 0;
  for (Py_ssize_t i = 0; i < n; ++i) {
    if (tk[i] <= 0) continue;                      /* nothing of this read is needed */
    if (PyObject_GetBuffer(PySequence_Fast_GET_ITEM(seq, i), &views[i], PyBUF_SIMPLE) != 0) goto done;
    got = i + 1;
    if (sk[i] < 0 || sk[i] + tk[i] > (int64_t)views[i].len || o[i] < 0 || o[i] + tk[i] > (int64_t)dst.len) {
      PyErr_Format(PyExc_ValueError, "pack: item %zd: range outside its buffer or the destination", i);
      goto done;
    }
    jobs[i].src = (const char*)views[i].buf + sk[i];
    jobs[i].dst = (char*)dst.buf + o[i];
    jobs[i].n = (size_t)tk[i];
    total += jobs[i].n;
  }
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 32) n_threads = 32;
  /* one thread per 8 MB: starting a thread costs about as much as copying a megabyte (measured on the GPU box:
     a 10 MB poll of 512 channels is fastest with 1-2 threads, a 60 MB poll of 3000 channels with 4) */
  if ((size_t)n_threads > 1 + (total >> 23)) n_threads = (int)(1 + (total >> 23));
  Py_BEGIN_ALLOW_THREADS
  if (n_threads == 1) {
    span_t s = {jobs, 0, n};
    run_span(&s);
  } else {
    /* contiguous spans of about equal bytes */
    span_t spans[32];
    pthread_t th[32];
    int started[32];
    size_t acc = 0;
    Py_ssize_t lo = 0;
    int t = 0;
    for (Py_ssize_t i = 0; i < n && t < n_threads - 1; ++i) {
      acc += jobs[i].n;
      if (acc >= total / (size_t)n_threads * (size_t)(t + 1)) {
        spans[t].jobs = jobs;
        spans[t].lo = lo;
        spans[t].hi = i + 1;
        lo = i + 1;
        ++t;
      }
    }
    spans[t].jobs = jobs;
    spans[t].lo = lo;
    spans[t].hi = n;
    ++t;
    for (int k = 1; k < t; ++k) started[k] = pthread_create(&th[k], NULL, run_span, &spans[k]) == 0;
    run_span(&spans[0]);
    for (int k = 1; k < t; ++k) {
      if (started[k]) pthread_join(th[k], NULL);
      else run_span(&spans[k]);
    }
  }
  Py_END_ALLOW_THREADS
  ret = Py_None;
  Py_INCREF(ret);
done:
  if (views) {
    for (Py_ssize_t i = 0; i < got; ++i)
      if (views[i].obj) PyBuffer_Release(&views[i]);
    free(views);
  }
  free(jobs);
  Py_XDECREF(seq);
  PyBuffer_Release(&dst);
  PyBuffer_Release(&off);
  PyBuffer_Release(&skip);
  PyBuffer_Release(&take);
  return ret;
}

static PyMethodDef methods[] = {
    {"lengths", hp_lengths, METH_VARARGS, "lengths(seq, out_int64): byte length of every buffer in seq"},
    {"pack", hp_pack, METH_VARARGS,
     "pack(seq, dst, byte_off, skip_bytes, take_bytes, n_threads=1): gather slices of the buffers in seq into dst"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef module = {PyModuleDef_HEAD_INIT, "_hostpack", "threaded gather of read prefixes", -1, methods};

PyMODINIT_FUNC PyInit__hostpack(void) { return PyModule_Create(&module); }
