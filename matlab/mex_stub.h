/* Declarations-only stand-in for MATLAB's mex.h so that matlab/fsae_mpc_b200_mex.c can be
 * syntax- and type-checked in an image without MATLAB (tests/test_cabi.py: gcc -fsyntax-only
 * -include mex_stub.h).  NOT a MEX runtime; never linked. */
#ifndef MEX_STUB_H
#define MEX_STUB_H
#include <stddef.h>
#include <stdio.h>
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
typedef enum { mxDOUBLE_CLASS = 6, mxINT8_CLASS = 8, mxINT32_CLASS = 12, mxUINT64_CLASS = 15 } mxClassID;
double* mxGetPr(const mxArray*);
void* mxGetData(const mxArray*);
double mxGetScalar(const mxArray*);
size_t mxGetM(const mxArray*);
size_t mxGetN(const mxArray*);
size_t mxGetNumberOfElements(const mxArray*);
size_t mxGetNumberOfDimensions(const mxArray*);
const mwSize* mxGetDimensions(const mxArray*);
int mxIsUint64(const mxArray*);
int mxIsInt32(const mxArray*);
int mxIsEmpty(const mxArray*);
int mxGetString(const mxArray*, char*, mwSize);
mxArray* mxCreateDoubleMatrix(mwSize, mwSize, mxComplexity);
mxArray* mxCreateNumericMatrix(mwSize, mwSize, mxClassID, mxComplexity);
mxArray* mxCreateNumericArray(mwSize, const mwSize*, mxClassID, mxComplexity);
void mexErrMsgTxt(const char*);
#endif
