// MOCK of the generated mpqc_config.h: the default build policy is sparse (CMakeLists.txt:61-70).
#pragma once
#define TA_DEFAULT_POLICY 1
