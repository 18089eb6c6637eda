/*
 * jdeflate/config/config.h -- static configuration of the B200-native build.
 *
 * Replaces the meson-generated file of the reference
 * (reference jdeflate/config/config.h.in:30-75): symbol visibility macro,
 * version macros, struct JDEFLATEVersion and jdeflate_getversion().
 * The version numbers track the reference release whose C API is mirrored
 * (reference VERSION: 0.4.0).
 */
#ifndef JDB200_JDEFLATE_CONFIG_H
#define JDB200_JDEFLATE_CONFIG_H

#if defined(_WIN32) && defined(JDEFLATE_DLL)
	#if defined(JDEFLATE_BUILDDLL)
		#define JDEFLATE_API __declspec(dllexport)
	#else
		#define JDEFLATE_API __declspec(dllimport)
	#endif
#elif defined(__GNUC__)
	#define JDEFLATE_API __attribute__((visibility("default")))
#else
	#define JDEFLATE_API
#endif

#define JDEFLATE_VERSION_MAJOR 0
#define JDEFLATE_VERSION_MINOR 4
#define JDEFLATE_VERSION_PATCH 0
#define JDEFLATE_VERSION_STRING "0.4.0"

/* marks this implementation; absent from the reference */
#define JDEFLATE_B200 1

#ifdef __cplusplus
extern "C" {
#endif

struct JDEFLATEVersion {
	int major;
	int minor;
	int patch;
	const char* versionstring;
	const char* builddate;
};

JDEFLATE_API
struct JDEFLATEVersion jdeflate_getversion(void);

#ifdef __cplusplus
}
#endif

#endif
