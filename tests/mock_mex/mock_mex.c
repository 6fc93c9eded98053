/* Implementation of the mock MEX runtime declared in tests/mock_mex/mex.h (TEST INFRASTRUCTURE). */
#include "mex.h"

#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static jmp_buf g_jmp;
static int g_in_call = 0;
static char g_err_id[128], g_err_msg[1024];
static struct { char name[64]; mxArray *val; } g_ws[64];
static int g_nws = 0, g_locks = 0;
static void (*g_atexit[8])(void);
static int g_natexit = 0;

mxArray *mxCreateDoubleMatrix(size_t m, size_t n, mxComplexity c) {
    mxArray *a = (mxArray *)calloc(1, sizeof(mxArray));
    a->m = m; a->n = n; a->is_double = 1; a->is_complex = (c == mxCOMPLEX);
    a->pr = (double *)calloc((m * n) > 0 ? m * n : 1, sizeof(double));
    return a;
}
mxArray *mxCreateDoubleScalar(double v) { mxArray *a = mxCreateDoubleMatrix(1, 1, mxREAL); a->pr[0] = v; return a; }
void mxDestroyArray(mxArray *a) { if (a) { free(a->pr); free(a); } }
double *mxGetPr(const mxArray *a) { return a->pr; }
double mxGetScalar(const mxArray *a) { return a->pr[0]; }
size_t mxGetM(const mxArray *a) { return a->m; }
size_t mxGetN(const mxArray *a) { return a->n; }
size_t mxGetNumberOfElements(const mxArray *a) { return a->m * a->n; }
int mxIsDouble(const mxArray *a) { return a->is_double; }
int mxIsComplex(const mxArray *a) { return a->is_complex; }
void *mxMalloc(size_t n) { return malloc(n ? n : 1); }
void mxFree(void *p) { free(p); }

void mexErrMsgIdAndTxt(const char *id, const char *fmt, ...) {
    va_list ap;
    snprintf(g_err_id, sizeof(g_err_id), "%s", id ? id : "");
    va_start(ap, fmt);
    vsnprintf(g_err_msg, sizeof(g_err_msg), fmt, ap);
    va_end(ap);
    if (g_in_call) longjmp(g_jmp, 1);
    fprintf(stderr, "mexErrMsgIdAndTxt outside mock_call_mex: %s: %s\n", g_err_id, g_err_msg);
    abort();
}
void mexWarnMsgIdAndTxt(const char *id, const char *fmt, ...) { (void)id; (void)fmt; }
const mxArray *mexGetVariablePtr(const char *workspace, const char *name) {
    int i;
    (void)workspace;
    for (i = 0; i < g_nws; ++i) if (!strcmp(g_ws[i].name, name)) return g_ws[i].val;
    return NULL;
}
void mexLock(void) { ++g_locks; }
int mexAtExit(void (*fn)(void)) { if (g_natexit < 8) g_atexit[g_natexit++] = fn; return 0; }
int mexPrintf(const char *fmt, ...) { va_list ap; int r; va_start(ap, fmt); r = vprintf(fmt, ap); va_end(ap); return r; }

int mock_call_mex(mex_entry_t fn, int nlhs, mxArray **plhs, int nrhs, const mxArray **prhs) {
    g_err_id[0] = g_err_msg[0] = 0;
    g_in_call = 1;
    if (setjmp(g_jmp)) { g_in_call = 0; return 1; }
    fn(nlhs, plhs, nrhs, prhs);
    g_in_call = 0;
    return 0;
}
const char *mock_last_error_id(void) { return g_err_id; }
const char *mock_last_error_msg(void) { return g_err_msg; }
void mock_set_variable(const char *name, mxArray *value) {
    int i;
    for (i = 0; i < g_nws; ++i) if (!strcmp(g_ws[i].name, name)) { g_ws[i].val = value; return; }
    if (g_nws < 64) { snprintf(g_ws[g_nws].name, sizeof(g_ws[g_nws].name), "%s", name); g_ws[g_nws++].val = value; }
}
void mock_clear_workspace(void) { g_nws = 0; }
int mock_lock_count(void) { return g_locks; }
void mock_run_atexit(void) { int i; for (i = g_natexit - 1; i >= 0; --i) g_atexit[i](); g_natexit = 0; }
