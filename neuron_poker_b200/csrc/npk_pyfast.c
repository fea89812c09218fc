/* npk_pyfast.c -- CPython binding of the one-query entry point npk_equity_one (include/npk.h) for the drop-in's hot call
 *
 *     get_equity(player_cards, table_cards, players, runs)            tools/montecarlo_python.py:401-406
 *     montecarlo(my_cards, cards_on_table, number_of_players, runs)   tools/montecarlo_cpp/pymontecarlo.cpp:21-23
 *
 * What it does is what neuron_poker_b200.equity.equity_counts does through ctypes -- parse the card strings, draw the call's
 * Philox seed from the caller's generator, call npk_equity_one, divide -- in C: a blocking call spends about 1 us here instead
 * of 3.5 us in Python byte code and ctypes argument conversion, which matters once the call itself is down to 9-17 us.
 * It only takes the common case: anything unusual (a card that is not a two-character string, a wrong number of cards, players
 * outside 1..10, runs <= 0, an error code from the library) returns None and the Python implementation runs, which raises the
 * reference's exceptions.  No computation happens here: the library is the product, this is a binding.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>

typedef int (*equity_one_fn)(uint64_t packed, int players, int64_t trials, uint64_t seed, int deal_mode, uint32_t want,
                             uint64_t* out);
static equity_one_fn g_equity_one = NULL;

/* 'AS' -> 4 * rank + suit (ranks "23456789TJQKA", suits "CDHS": MonteCarlo.create_card_deck, montecarlo_python.py:114-119) */
static int card_id(PyObject* s)
{
    if (!PyUnicode_CheckExact(s) || PyUnicode_GET_LENGTH(s) != 2 || PyUnicode_KIND(s) != PyUnicode_1BYTE_KIND) return -1;
    const Py_UCS1* d = PyUnicode_1BYTE_DATA(s);
    int r, u;
    switch (d[0]) {
        case '2': r = 0; break; case '3': r = 1; break; case '4': r = 2; break; case '5': r = 3; break;
        case '6': r = 4; break; case '7': r = 5; break; case '8': r = 6; break; case '9': r = 7; break;
        case 'T': r = 8; break; case 'J': r = 9; break; case 'Q': r = 10; break; case 'K': r = 11; break;
        case 'A': r = 12; break; default: return -1;
    }
    switch (d[1]) {
        case 'C': u = 0; break; case 'D': u = 1; break; case 'H': u = 2; break; case 'S': u = 3; break;
        default: return -1;
    }
    return 4 * r + u;
}

/* card ids of a set / list / tuple of at most `max` cards into ids[] (iteration order, like the Python implementation);
 * returns the count or -1 for "not the common case" (no exception left pending) */
static int collect(PyObject* cards, int* ids, int max)
{
    int n = 0;
    if (PyList_CheckExact(cards) || PyTuple_CheckExact(cards)) {
        const Py_ssize_t len = PySequence_Fast_GET_SIZE(cards);
        if (len > max) return -1;
        for (Py_ssize_t i = 0; i < len; i++) {
            const int c = card_id(PySequence_Fast_GET_ITEM(cards, i));
            if (c < 0) return -1;
            ids[n++] = c;
        }
        return n;
    }
    if (!PyAnySet_CheckExact(cards) || PySet_GET_SIZE(cards) > max) return -1;
    PyObject* it = PyObject_GetIter(cards);
    if (!it) { PyErr_Clear(); return -1; }
    PyObject* item;
    while ((item = PyIter_Next(it)) != NULL) {
        const int c = card_id(item);
        Py_DECREF(item);
        if (c < 0 || n == max) { Py_DECREF(it); return -1; }
        ids[n++] = c;
    }
    Py_DECREF(it);
    if (PyErr_Occurred()) { PyErr_Clear(); return -1; }
    return n;
}

static long long as_index(PyObject* o)          /* int, numpy integer, ...; -1 when it is not an integer */
{
    if (PyLong_CheckExact(o)) return PyLong_AsLongLong(o);
    PyObject* i = PyNumber_Index(o);
    if (!i) { PyErr_Clear(); return -1; }
    const long long v = PyLong_AsLongLong(i);
    Py_DECREF(i);
    return v;
}

/* equity(player_cards, table_cards, players, runs, deal_mode, seed_fn) -> float | None
 * seed_fn() returns a float in [0, 1) (numpy's legacy random_sample: the generator the reference itself consumes) */
static PyObject* equity(PyObject* self, PyObject* const* args, Py_ssize_t nargs)
{
    (void)self;
    if (nargs != 6 || !g_equity_one) Py_RETURN_NONE;
    int hole[2], board[5];
    if (collect(args[0], hole, 2) != 2) Py_RETURN_NONE;
    const int nb = collect(args[1], board, 5);
    if (nb < 0) Py_RETURN_NONE;
    const long long players = as_index(args[2]);
    const long long runs = as_index(args[3]);
    const long mode = PyLong_AsLong(args[4]);
    if (PyErr_Occurred()) { PyErr_Clear(); Py_RETURN_NONE; }
    if (players < 1 || players > 10 || runs <= 0) Py_RETURN_NONE;
    PyObject* r = PyObject_CallNoArgs(args[5]);
    if (!r) return NULL;
    const double u = PyFloat_AsDouble(r);
    Py_DECREF(r);
    if (u == -1.0 && PyErr_Occurred()) return NULL;
    const uint64_t seed = (uint64_t)(u * 9007199254740992.0);
    uint64_t packed = (uint64_t)hole[0] | (uint64_t)hole[1] << 8;
    for (int i = 0; i < 5; i++) packed |= (uint64_t)(i < nb ? board[i] : 0xFF) << (16 + 8 * i);
    uint64_t out[12];
    int rc;
    Py_BEGIN_ALLOW_THREADS
    rc = g_equity_one(packed, (int)players, (int64_t)runs, seed, (int)mode, 0u, out);
    Py_END_ALLOW_THREADS
    if (rc != 0) Py_RETURN_NONE;              /* the Python path repeats the call and raises the library's error */
    return PyFloat_FromDouble((double)(out[0] + out[1]) / (double)runs);
}

/* bind(address of npk_equity_one in the libnpk.so the package has loaded) */
static PyObject* bind(PyObject* self, PyObject* arg)
{
    (void)self;
    void* p = PyLong_AsVoidPtr(arg);
    if (!p && PyErr_Occurred()) return NULL;
    g_equity_one = (equity_one_fn)p;
    Py_RETURN_NONE;
}

static PyMethodDef methods[] = {
    {"equity", (PyCFunction)(void (*)(void))equity, METH_FASTCALL, "one get_equity call through npk_equity_one, or None"},
    {"bind", bind, METH_O, "set the address of npk_equity_one"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef module = {PyModuleDef_HEAD_INIT, "_npkfast", "C binding of npk_equity_one", -1, methods,
                                    NULL, NULL, NULL, NULL};

PyMODINIT_FUNC PyInit__npkfast(void) { return PyModule_Create(&module); }
