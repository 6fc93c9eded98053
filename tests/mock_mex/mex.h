/* Minimal stand-in for MATLAB's mex.h (TEST INFRASTRUCTURE): just enough of the real-double matrix API for the
 * gateways in mpc-ntm-control_b200/csrc/mex/ to compile and run without MATLAB/Octave (neither exists in the
 * build image).  Semantics follow the documented MEX contract: column-major doubles, prhs borrowed,
 * plhs owned by the caller after return, mexErrMsgIdAndTxt never returns (here: longjmp to mock_call_mex). */
#ifndef MOCK_MEX_H
#define MOCK_MEX_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct mxArray_tag { size_t m, n; double *pr; int is_double; int is_complex; } mxArray;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;

mxArray *mxCreateDoubleMatrix(size_t m, size_t n, mxComplexity c);
mxArray *mxCreateDoubleScalar(double v);
void mxDestroyArray(mxArray *a);
double *mxGetPr(const mxArray *a);
double mxGetScalar(const mxArray *a);
size_t mxGetM(const mxArray *a);
size_t mxGetN(const mxArray *a);
size_t mxGetNumberOfElements(const mxArray *a);
int mxIsDouble(const mxArray *a);
int mxIsComplex(const mxArray *a);
void *mxMalloc(size_t n);
void mxFree(void *p);
void mexErrMsgIdAndTxt(const char *id, const char *fmt, ...);
void mexWarnMsgIdAndTxt(const char *id, const char *fmt, ...);
const mxArray *mexGetVariablePtr(const char *workspace, const char *name);
void mexLock(void);
int mexAtExit(void (*fn)(void));
int mexPrintf(const char *fmt, ...);

/* --- mock control surface (used by tests/test_mex_gateway.py through ctypes) --- */
typedef void (*mex_entry_t)(int, mxArray **, int, const mxArray **);
int mock_call_mex(mex_entry_t fn, int nlhs, mxArray **plhs, int nrhs, const mxArray **prhs); /* 0 ok, 1 error raised */
const char *mock_last_error_id(void);
const char *mock_last_error_msg(void);
void mock_set_variable(const char *name, mxArray *value);   /* caller workspace; NULL value clears */
void mock_clear_workspace(void);
int mock_lock_count(void);
void mock_run_atexit(void);
#ifdef __cplusplus
}
#endif
#endif
