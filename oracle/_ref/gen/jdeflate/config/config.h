 /*
 * Copyright (C) 2025, jpn
 *
 * Licensed under the Apache License, Version 2.0 (the "License");
 * you may not use this file except in compliance with the License.
 * You may obtain a copy of the License at
 *
 * http://www.apache.org/licenses/LICENSE-2.0
 *
 * Unless required by applicable law or agreed to in writing, software
 * distributed under the License is distributed on an "AS IS" BASIS,
 * WITHOUT WARRANTIES OR CONDITIONS OF ANY KIND, either express or implied.
 * See the License for the specific language governing permissions and
 * limitations under the License.
 */

#ifndef d5c77bd8_2a38_4e2d_a74f_c53348f5e61e
#define d5c77bd8_2a38_4e2d_a74f_c53348f5e61e

/*
 * config.h
 * Configuration flags.
 */


/* ***************************************************************************
 * Function visibility attributes
 *************************************************************************** */

#if defined(JDEFLATE_DLL)
	#if defined(_WIN32)
		#if defined(JDEFLATE_BUILDDLL)
			#define JDEFLATE_API __declspec(dllexport)
		#else
			#define JDEFLATE_API __declspec(dllimport)
		#endif
	#endif
#endif

#ifndef JDEFLATE_API
	#if defined(__GNUC__)
		#define JDEFLATE_API __attribute__((visibility("default")))
	#else
		#define JDEFLATE_API
	#endif
#endif


/* ***************************************************************************
 * Version number
 *************************************************************************** */

#define JDEFLATE_VERSION_MAJOR 0
#define JDEFLATE_VERSION_MINOR 4
/* revision */
#define JDEFLATE_VERSION_PATCH 0


#define JDEFLATE_VERSION_STRING "0.4.0"


/*
 * Version info */
struct JDEFLATEVersion {
	int major;
	int minor;
	int patch;
	const char* versionstring;
	const char* builddate;
};

/*
 * */
JDEFLATE_API
struct JDEFLATEVersion jdeflate_getversion(void);


#endif
