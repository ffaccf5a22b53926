/* stand-in for VTK's vtkStreamingDemandDrivenPipeline.h: see vtk_standin.h (test infrastructure) */
#include "vtk_standin.h"
