// Oracle build shim: included by clustering/ReadClusteringEngine.h:3, nothing from it is used.
#pragma once
