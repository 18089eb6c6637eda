/* version.c -- jdeflate_getversion() (reference src/version.c:20-35) */
#include <jdeflate/config/config.h>

static const char jdb_version_text[] = JDEFLATE_VERSION_STRING;

struct JDEFLATEVersion
jdeflate_getversion(void)
{
	struct JDEFLATEVersion info;

	info.versionstring = jdb_version_text;
	info.builddate = (const char*) 0;
	info.major = JDEFLATE_VERSION_MAJOR;
	info.minor = JDEFLATE_VERSION_MINOR;
	info.patch = JDEFLATE_VERSION_PATCH;
	return info;
}
