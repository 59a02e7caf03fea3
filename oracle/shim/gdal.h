// TEST INFRASTRUCTURE ONLY (oracle build).  The subset of the GDAL/OGR C API
// that the reference's hot-path sources call, backed by
// the repo's native TIFF reader/writer.  It moves bytes and converts types; it
// computes nothing.
#pragma once
#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef __cplusplus
#include <string>
extern "C" {
#endif
#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif
typedef void* GDALDatasetH;
typedef void* GDALDriverH;
typedef void* GDALRasterBandH;
typedef void* OGRSpatialReferenceH;
typedef void* OGRDataSourceH;
typedef void* OGRSFDriverH;
typedef void* OGRLayerH;
typedef void* OGRFeatureDefnH;
typedef void* OGRFieldDefnH;
typedef void* OGRFeatureH;
typedef void* OGRGeometryH;
typedef int CPLErr;
typedef int OGRErr;
typedef enum { GDT_Unknown = 0, GDT_Byte = 1, GDT_UInt16 = 2, GDT_Int16 = 3, GDT_UInt32 = 4, GDT_Int32 = 5, GDT_Float32 = 6, GDT_Float64 = 7 } GDALDataType;
typedef enum { GA_ReadOnly = 0, GA_Update = 1 } GDALAccess;
typedef enum { GF_Read = 0, GF_Write = 1 } GDALRWFlag;
typedef enum { wkbUnknown = 0, wkbPoint = 1, wkbLineString = 2, wkbPolygon = 3, wkbMultiPoint = 4 } OGRwkbGeometryType;
typedef enum { OFTInteger = 0, OFTIntegerList = 1, OFTReal = 2, OFTRealList = 3, OFTString = 4, OFTInteger64 = 12 } OGRFieldType;
typedef long long GIntBig;

void GDALAllRegister(void);
GDALDatasetH GDALOpen(const char*, GDALAccess);
void GDALClose(GDALDatasetH);
void GDALFlushCache(GDALDatasetH);
GDALDriverH GDALGetDatasetDriver(GDALDatasetH);
GDALDriverH GDALGetDriverByName(const char*);
GDALDatasetH GDALCreate(GDALDriverH, const char*, int, int, int, GDALDataType, char**);
const char* GDALGetProjectionRef(GDALDatasetH);
CPLErr GDALSetProjection(GDALDatasetH, const char*);
CPLErr GDALGetGeoTransform(GDALDatasetH, double*);
CPLErr GDALSetGeoTransform(GDALDatasetH, double*);
GDALRasterBandH GDALGetRasterBand(GDALDatasetH, int);
int GDALGetRasterXSize(GDALDatasetH);
int GDALGetRasterYSize(GDALDatasetH);
const char* GDALGetRasterUnitType(GDALRasterBandH);
GDALDataType GDALGetRasterDataType(GDALRasterBandH);
double GDALGetRasterNoDataValue(GDALRasterBandH, int*);
CPLErr GDALSetRasterNoDataValue(GDALRasterBandH, double);
CPLErr GDALRasterIO(GDALRasterBandH, GDALRWFlag, int, int, int, int, void*, int, int, GDALDataType, int, int);
char** CSLSetNameValue(char**, const char*, const char*);
const char* CPLGetLastErrorMsg(void);

OGRSpatialReferenceH OSRNewSpatialReference(const char*);
int OSRIsGeographic(OGRSpatialReferenceH);
int OSRIsProjected(OGRSpatialReferenceH);
double OSRGetLinearUnits(OGRSpatialReferenceH, char**);
const char* OSRGetAttrValue(OGRSpatialReferenceH, const char*, int);

void OGRRegisterAll(void);
OGRDataSourceH OGROpen(const char*, int, OGRSFDriverH*);
OGRLayerH OGR_DS_GetLayer(OGRDataSourceH, int);
OGRLayerH OGR_DS_GetLayerByName(OGRDataSourceH, const char*);
int OGR_DS_GetLayerCount(OGRDataSourceH);
void OGR_DS_Destroy(OGRDataSourceH);
const char* OGR_L_GetName(OGRLayerH);
OGRwkbGeometryType OGR_L_GetGeomType(OGRLayerH);
OGRSpatialReferenceH OGR_L_GetSpatialRef(OGRLayerH);
GIntBig OGR_L_GetFeatureCount(OGRLayerH, int);
OGRFeatureDefnH OGR_L_GetLayerDefn(OGRLayerH);
void OGR_L_ResetReading(OGRLayerH);
OGRFeatureH OGR_L_GetNextFeature(OGRLayerH);
OGRFeatureH OGR_L_GetFeature(OGRLayerH, GIntBig);
OGRGeometryH OGR_F_GetGeometryRef(OGRFeatureH);
int OGR_F_GetFieldIndex(OGRFeatureH, const char*);
int OGR_F_GetFieldAsInteger(OGRFeatureH, int);
GIntBig OGR_F_GetFieldAsInteger64(OGRFeatureH, int);
double OGR_F_GetFieldAsDouble(OGRFeatureH, int);
const char* OGR_F_GetFieldAsString(OGRFeatureH, int);
void OGR_F_Destroy(OGRFeatureH);
OGRFieldDefnH OGR_FD_GetFieldDefn(OGRFeatureDefnH, int);
OGRFieldType OGR_Fld_GetType(OGRFieldDefnH);
double OGR_G_GetX(OGRGeometryH, int);
double OGR_G_GetY(OGRGeometryH, int);
#ifdef __cplusplus
}
#endif
