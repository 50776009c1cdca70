/* Stub <windows.h> for compiling the reference's numeric headers on Linux.
 * Test infrastructure only (oracle/_ref build); see oracle/Makefile.
 * The reference's Lock.h (M/Lock.h:5) only needs a CRITICAL_SECTION type and
 * four functions; the oracle build is single-threaded so they are no-ops. */
#ifndef UAVM_REF_SHIM_WINDOWS_H
#define UAVM_REF_SHIM_WINDOWS_H
typedef struct { int unused; } CRITICAL_SECTION;
static inline void InitializeCriticalSection(CRITICAL_SECTION*) {}
static inline void DeleteCriticalSection(CRITICAL_SECTION*) {}
static inline void EnterCriticalSection(CRITICAL_SECTION*) {}
static inline void LeaveCriticalSection(CRITICAL_SECTION*) {}
typedef unsigned long DWORD;
typedef void* LPVOID;
typedef void* HANDLE;
#define WINAPI
#endif
